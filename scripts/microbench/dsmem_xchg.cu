// Microbenchmark of the split-K exchange inside a cluster: every CTA holds a 128 x NB fp32 partial tile (one row per
// thread of 4 warps) and ships rows [o*RO, (o+1)*RO) to CTA o of the cluster (RO = 128 / C), which needs all C blocks.
//   variant 0: STS into a staging tile + fence.proxy.async + one cp.async.bulk (DSMEM) per destination  (round-1 design)
//   variant 1: st.async.v4 straight from registers into the owner's slot (mbarrier complete_tx per 16 bytes)
//   variant 2: plain st.shared::cluster.v4 + remote mbarrier arrive (release.cluster) per warp
// Reports cycles per exchange round (128 rounds, slots ping-ponged, a cluster barrier every round keeps the CTAs in
// lockstep; the barrier-only loop is timed as the baseline to subtract).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include -I../../dr-nmf_b200/csrc dsmem_xchg.cu -o dsmem_xchg
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace drnmf;

namespace drnmf { void set_error(const char*, ...) {} }

template <int NB, int VAR>
__global__ void __launch_bounds__(256, 1) k_xchg(long long* out, int rounds, int do_xfer) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[2];
  const int C = gridDim.x;               // cluster = whole grid.x
  const int RO = 128 / C;
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* stage = smem;                         // 128 x NB x 4
  uint8_t* red = smem + 128 * NB * 4;            // 2 slots x (C blocks of RO x NB x 4) = 2 x 128 x NB x 4
  const uint32_t slot_bytes = 128 * NB * 4, blk_bytes = RO * NB * 4;
  if (threadIdx.x == 0) {
    // variant 2 counts one release-arrive per source CTA (the rows of an owner lie inside one warp: RO <= 32);
    // the others count bytes (one expect_tx arrival)
    mbar_init(&full[0], VAR == 2 ? C : 1); mbar_init(&full[1], VAR == 2 ? C : 1);
    fence_mbar_init();
  }
  __syncthreads();
  cluster_sync_all();
  float v[NB];
  for (int c = 0; c < NB; ++c) v[c] = (float)(threadIdx.x + c);
  long long t0 = clock64();
  float sink = 0.f;
  for (int r = 0; r < rounds; ++r) {
    const int sl = r & 1;
    if (do_xfer) {
      if (warp < 4) {                              // pushers: thread = row
        const int rho = threadIdx.x;
        const uint32_t o = rho / RO;
        if (VAR == 0) {
          const uint32_t srow = smem_u32(stage) + rho * (NB * 4);
#pragma unroll
          for (int c = 0; c < NB / 4; ++c) {
            const uint32_t addr = srow + (uint32_t)(((c & ~7) | ((c ^ rho) & 7)) * 16);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * c]), "f"(v[4 * c + 1]), "f"(v[4 * c + 2]), "f"(v[4 * c + 3]) : "memory");
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (warp == 0 && lane < C) {
            const uint32_t oo = lane;
            const uint32_t src = smem_u32(stage) + oo * blk_bytes;
            const uint32_t dst = mapa_u32(smem_u32(red) + sl * slot_bytes + rank * blk_bytes, oo);
            const uint32_t bar = mapa_u32(smem_u32(&full[sl]), oo);
            dsmem_bulk_copy(dst, src, blk_bytes, bar);
          }
        } else if (VAR == 1) {
          const uint32_t dst = mapa_u32(smem_u32(red) + sl * slot_bytes + rank * blk_bytes + (rho % RO) * (NB * 4), o);
          const uint32_t bar = mapa_u32(smem_u32(&full[sl]), o);
#pragma unroll
          for (int c = 0; c < NB / 4; ++c) st_async_v4(dst + c * 16, bar, v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        } else {
          const uint32_t dst = mapa_u32(smem_u32(red) + sl * slot_bytes + rank * blk_bytes + (rho % RO) * (NB * 4), o);
#pragma unroll
          for (int c = 0; c < NB / 4; ++c)
            asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + c * 16), "f"(v[4 * c]), "f"(v[4 * c + 1]), "f"(v[4 * c + 2]), "f"(v[4 * c + 3]) : "memory");
          __syncwarp();
          // a warp's 32 rows belong to 32/RO owners: one release-arrive per (warp, owner)
          if (lane % RO == 0) mbar_arrive_remote(&full[sl], o);
        }
      } else {                                     // owners (warps 4..7): wait for all C blocks, touch them
        if (VAR != 2 && threadIdx.x == 128) mbar_expect_tx(&full[sl], slot_bytes);
        mbar_wait_cluster(&full[sl], (r >> 1) & 1);
        const float* p = reinterpret_cast<const float*>(red + sl * slot_bytes);
        sink += p[(threadIdx.x - 128) * 4];
      }
    }
    cluster_sync_all();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / rounds;
  if (sink == 123.456f) out[1] = 1;
}

template <int NB, int VAR>
static void run(int C, const char* name, long long* d_out) {
  const int smem = 3 * 128 * NB * 4 + 1024;
  cudaFuncSetAttribute(k_xchg<NB, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long res[2][2];
  for (int x = 0; x < 2; ++x) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(C, 1, 1); cfg.blockDim = dim3(256, 1, 1); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_xchg<NB, VAR>, d_out, 128, x);
    if (e != cudaSuccess) { printf("%s launch failed: %s\n", name, cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(res[x], d_out, 16, cudaMemcpyDeviceToHost);
  }
  printf("C=%d NB=%2d %-10s: %6lld cycles/round (barrier-only %lld) -> exchange %lld\n", C, NB, name, res[1][0], res[0][0], res[1][0] - res[0][0]);
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 64);
  for (int C : {4, 8}) {
    run<16, 0>(C, "bulk", d_out); run<16, 1>(C, "st.async", d_out); run<16, 2>(C, "st+arrive", d_out);
    run<32, 0>(C, "bulk", d_out); run<32, 1>(C, "st.async", d_out); run<32, 2>(C, "st+arrive", d_out);
    run<64, 0>(C, "bulk", d_out); run<64, 1>(C, "st.async", d_out); run<64, 2>(C, "st+arrive", d_out);
  }
  return 0;
}
