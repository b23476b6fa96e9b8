// Microbenchmark for the product phase of the persistent recurrence: a chain of tcgen05.mma.kind::tf32 with A in TMEM
// (M = 128) and B in shared memory (K-major, 128B swizzle), all accumulating into ONE TMEM accumulator, as issued by
// warp 2 of k_recurrent_tc.  Prints cycles from the first issue to the completion of the commit.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include -I../../dr-nmf_b200/csrc mma_rate.cu ../../dr-nmf_b200/csrc/runtime.cu -o mma_rate -lcuda
#include <cstdio>
#include <cuda.h>
#include "common.cuh"
using namespace drnmf;

template <int N, int NMMA, int NACC, bool SINGLE_THREAD, int PATTERN = 0, int WAITERS = 0>
__global__ void __launch_bounds__(512, 1) k_rate(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, park;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&park, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.5f;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  {   // A operand: 256 columns of ones in every lane
    float v[32];
    for (int e = 0; e < 32; ++e) v[e] = 1.0f;
    if (warp < 4) for (int c = 0; c < 256; c += 32) tmem_st32(tb + ((uint32_t)(warp * 32) << 16) + c, v);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, N);
    const uint32_t hb = smem_u32(smem);
    const uint64_t d0 = umma_desc_k128(hb);
    for (int rep = 0; rep < 6; ++rep) {
      const bool leader = SINGLE_THREAD ? (lane == 0) : elect_one();
      __syncwarp();
      const long long t0 = clock64();
      if (!SINGLE_THREAD || lane == 0) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i) {
          uint32_t a = tb + (i % 32) * 8;
          // B tiles of N rows x 128 bytes; 4 k-steps of 32 bytes inside a tile.  Descriptor address field is in 16-byte units.
          uint64_t b = d0 + (uint64_t)((((i / 4) % 8) * (N * 128) + (i % 4) * 32) >> 4);
          if (PATTERN != 0) {   // k-step ks = i / 3: slab at = ks / 4, kk = ks % 4; tiles [2 at] = hi, [2 at + 1] = lo; A hi at col ks*8, lo at 128 + ks*8
            const int ks = (i / 3) % 16, at = ks / 4, kk = ks % 4, j = i % 3;
            const bool a_lo = (PATTERN == 1) ? (j == 0) : (j == 2);
            const bool b_lo = (j == 1);
            a = tb + (a_lo ? 128 : 0) + ks * 8;
            b = d0 + (uint64_t)((((2 * at + (b_lo ? 1 : 0)) % 8) * (N * 128) + kk * 32) >> 4);
          }
          if (leader) umma_tf32_ts(tb + 256 + (i % NACC) * N, a, b, idesc, i >= NACC);
        }
      }
      const long long t1 = clock64();
      if (leader) tc_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      if (threadIdx.x == 0 && blockIdx.x == 0) { out[2 * rep] = t1 - t0; out[2 * rep + 1] = t2 - t0; }
      if (lane == 0) mbar_arrive(&park);     // releases the parked warps of this repetition
    }
  } else if (warp <= WAITERS) {
    for (int rep = 0; rep < 6; ++rep) mbar_wait(&park, rep & 1);   // all 32 lanes wait, like the roles of the recurrence
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tb); }
}

template <int N, int NMMA, int NACC, bool ST, int PAT = 0, int WAITERS = 0>
static void run(long long* d, const char* what, int grid = 1) {
  auto k = k_rate<N, NMMA, NACC, ST, PAT, WAITERS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
  if (grid < 0) {   // cluster launch: -grid CTAs in clusters of 8
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(-grid, 1, 1); cfg.blockDim = dim3(32 * (WAITERS + 1) < 128 ? 128 : 32 * (WAITERS + 1), 1, 1);
    cfg.dynamicSmemBytes = 131072 + 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k, d);
  } else
  k<<<grid, 32 * (WAITERS + 1) < 128 ? 128 : 32 * (WAITERS + 1), 131072 + 1024>>>(d);
  long long h[12];
  if (cudaMemcpy(h, d, 96, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("%s: error %s\n", what, cudaGetErrorString(cudaGetLastError())); exit(1); }
  printf("grid=%3d N=%3d n_mma=%2d acc=%d %-14s issue %5lld  complete %5lld cycles  = %5.1f per MMA, %6.0f MAC/clk\n", grid, N, NMMA, NACC, what, h[10], h[11],
         (double)h[11] / NMMA, 128.0 * N * 8 * NMMA / (double)h[11]);
}

int main() {
  long long* d; cudaMalloc(&d, 96);
  run<64, 48, 1, false>(d, "distinct");
  run<64, 48, 1, false, 1>(d, "lo.hi,hi.lo,hi.hi");
  run<64, 48, 1, false, 2>(d, "hi.hi,hi.lo,lo.hi");
  run<64, 48, 1, false>(d, "distinct", 148);
  run<64, 48, 1, false, 1, 3>(d, "3 parked warps");
  run<64, 48, 1, false, 1, 7>(d, "7 parked warps");
  run<64, 48, 1, false, 1, 15>(d, "15 parked warps");
  run<64, 48, 1, false, 1, 15>(d, "cluster 8x8", -64);
  run<64, 16, 1, false>(d, "elect");
  run<64, 96, 1, false>(d, "elect");
  run<64, 48, 1, true>(d, "single thread");
  run<128, 24, 1, false>(d, "elect");
  run<32, 48, 1, false>(d, "elect");
  run<16, 48, 1, false>(d, "elect");
  return 0;
}
