// C = A . B^T with fused epilogues, in two implementations over identical layouts:
//   k_gemm_tc   : tcgen05.mma kind::tf32 with TMEM accumulators, operands staged by TMA (128B swizzle), 3xTF32
//                 error compensation (A_lo.B_hi + A_hi.B_lo + A_hi.B_hi) so that results keep fp32 fidelity.
//   k_gemm_simt : CUDA-core fp32 FMAs (semantics lock / debugging).
// Used for: the Gram build S_k (enhance.py:172-181), the input projections x~.W_k (custom_layers.py:368 hoisted out
// of the recurrence) and recon + mask (enhance.py:269-305, custom_layers.py:24,44).
#include "internal.h"
#include "gemm_simt.cuh"

#include <cstdlib>

namespace drnmf {

// ------------------------------------------------------------------------------------------------
// SIMT
// ------------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(SIMT_THREADS) k_gemm_simt(GemmArgs a) {
  float acc[4][4], acc2[4][4];
  const int m0 = blockIdx.y * SIMT_BM, n0 = blockIdx.x * SIMT_BN;
  const int nsplit = a.splits > 1 ? a.splits : 1;
  const int kchunk = ((a.Kd + nsplit - 1) / nsplit + SIMT_BK - 1) / SIMT_BK * SIMT_BK;
  const int zsplit = blockIdx.z + a.split_z0;
  const int kbeg = zsplit * kchunk;
  int klen = a.Kd - kbeg; if (klen > kchunk) klen = kchunk; if (klen < 0) klen = 0;
  constexpr bool DUAL = (EPI == EPI_RECON || EPI == EPI_MU_H);
  simt_tile_mainloop<DUAL>(a.A_hi + kbeg, a.lda, a.M, a.B_hi + kbeg, DUAL ? a.B2_hi + kbeg : nullptr,
                                       a.ldb, a.N, klen, m0, n0, acc, acc2);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float* Cz = a.C + (size_t)zsplit * a.split_stride;
  float dsum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.M_valid) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.N_valid) continue;
      if (EPI == EPI_STORE) {
        Cz[(size_t)m * a.ldc + n] = acc[i][j] + (a.bias ? a.bias[n] : 0.f);
      } else if (EPI == EPI_GRAM) {
        float v = epi_gram_value(m, n, a.R_valid, acc[i][j]);
        a.C[(size_t)m * a.ldc + n] = v;
        a.C_lo[(size_t)m * a.ldc + n] = tf32_lo(v);
      } else if (EPI == EPI_RECON) {
        a.C[(size_t)m * a.ldc + n] = epi_irm_value(acc[i][j], acc2[i][j], a.square);
        if (a.C_S) { a.C_S[(size_t)m * a.ldc + n] = acc[i][j]; a.C_N[(size_t)m * a.ldc + n] = acc2[i][j]; }
      } else if (EPI == EPI_MU_H) {
        float hv = a.C[(size_t)m * a.ldc + n];
        if (!a.row_update || a.row_update[m]) hv = hv * acc2[i][j] / fmaxf(acc[i][j] + a.mu, a.flr);
        a.C[(size_t)m * a.ldc + n] = hv; a.C_lo[(size_t)m * a.ldc + n] = tf32_lo(hv);
        a.CT[(size_t)n * a.ldct + m] = hv; a.CT_lo[(size_t)n * a.ldct + m] = tf32_lo(hv);
        dsum += hv;
      } else if (EPI == EPI_LAMBDA) {
        const float v = fmaxf(acc[i][j], a.flr);
        a.C[(size_t)m * a.ldc + n] = v; a.C_lo[(size_t)m * a.ldc + n] = tf32_lo(v);
        a.CT[(size_t)n * a.ldct + m] = v; a.CT_lo[(size_t)n * a.ldct + m] = tf32_lo(v);
        const float d = a.Vref[(size_t)m * a.ldv + n] - v;
        dsum = fmaf(d, d, dsum);
      } else {
        float P, Q, d;
        epi_beta_values(fmaxf(acc[i][j], a.flr), a.Vref[(size_t)m * a.ldv + n], a.beta, P, Q, d);
        a.C[(size_t)m * a.ldc + n] = P; a.C_lo[(size_t)m * a.ldc + n] = tf32_lo(P);
        a.CT[(size_t)n * a.ldct + m] = P; a.CT_lo[(size_t)n * a.ldct + m] = tf32_lo(P);
        a.Q[(size_t)m * a.ldc + n] = Q; a.Q_lo[(size_t)m * a.ldc + n] = tf32_lo(Q);
        a.QT[(size_t)n * a.ldct + m] = Q; a.QT_lo[(size_t)n * a.ldct + m] = tf32_lo(Q);
        dsum += d;
      }
    }
  }
  if (EPI == EPI_LAMBDA || EPI == EPI_LAMBDA_B || EPI == EPI_MU_H) {
    __shared__ float red[SIMT_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < SIMT_THREADS / 32; ++w) t += (double)red[w];
      a.div_partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
  }
}

int launch_gemm_simt(GemmEpi epi, const GemmArgs& a, cudaStream_t st) {
  dim3 grid((a.N + SIMT_BN - 1) / SIMT_BN, (a.M + SIMT_BM - 1) / SIMT_BM, a.split_nz > 0 ? a.split_nz : (a.splits > 1 ? a.splits : 1));
  switch (epi) {
    case EPI_STORE: k_gemm_simt<EPI_STORE><<<grid, SIMT_THREADS, 0, st>>>(a); break;
    case EPI_GRAM:  k_gemm_simt<EPI_GRAM><<<grid, SIMT_THREADS, 0, st>>>(a); break;
    case EPI_RECON: k_gemm_simt<EPI_RECON><<<grid, SIMT_THREADS, 0, st>>>(a); break;
    case EPI_LAMBDA: k_gemm_simt<EPI_LAMBDA><<<grid, SIMT_THREADS, 0, st>>>(a); break;
    case EPI_LAMBDA_B: k_gemm_simt<EPI_LAMBDA_B><<<grid, SIMT_THREADS, 0, st>>>(a); break;
    case EPI_MU_H: k_gemm_simt<EPI_MU_H><<<grid, SIMT_THREADS, 0, st>>>(a); break;
  }
  count_launch();
  DRNMF_CUDA(cudaGetLastError());
  return DRNMF_OK;
}

// ------------------------------------------------------------------------------------------------
// tcgen05
// ------------------------------------------------------------------------------------------------
constexpr int TC_BM = 128, TC_BK = 32;                        // BK = one 128-byte swizzle atom of fp32
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;             // 16 KB per A tile (and per 128-row B tile)
constexpr int TC_THREADS = 192;                              // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue

// BN = 128 or 256 output columns per CTA.  With both operands in shared memory an M128 x N x K8 TF32 MMA reads
// (128 + N) * 32 bytes per N/2 cycles: 128 B/clk at N = 128 (the whole shared-memory port, while TMA refills the ring),
// 96 B/clk at N = 256 - the wide tile is used whenever the grid still fills the device.
template <int EPI, int BN> struct TcCfg {
  static constexpr bool DUAL = (EPI == EPI_RECON || EPI == EPI_MU_H);
  static_assert(!(DUAL && BN != 128), "the dual-B epilogue needs both accumulators: 128 columns each");
  static constexpr int B_TILE_BYTES = BN * TC_BK * 4;
  static constexpr int STAGES = (DUAL || BN == 256) ? 2 : 3;
  static constexpr int STAGE_BYTES = 2 * TC_TILE_BYTES + (DUAL ? 4 : 2) * B_TILE_BYTES;   // A_hi A_lo B_hi B_lo [B2_hi B2_lo]
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  // tcgen05 accumulates with round-toward-zero: a long chain of MMAs into one accumulator is biased low by about half an
  // ulp per instruction.  The k-blocks are therefore dealt round-robin to NACC independent TMEM accumulators (all 512
  // columns) that the epilogue adds in fp32 round-to-nearest: chains are NACC times shorter, so is the bias.
  static constexpr int NACC = (DUAL || BN == 256) ? 2 : 4;
  static constexpr uint32_t ACC_COLS = DUAL ? 256 : BN;
  static constexpr uint32_t TMEM_COLS = 512;
};

template <int EPI, int TC_BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
          const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
          const __grid_constant__ CUtensorMap tmB2_hi, const __grid_constant__ CUtensorMap tmB2_lo, GemmArgs a,
          int* dev_error) {
  using Cfg = TcCfg<EPI, TC_BN>;
  extern __shared__ __align__(1024) uint8_t smem[];      // no static smem: the dynamic window is 1024-aligned
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* acc_full = empty + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  if ((smem_u32(smem) & 1023u) != 0) { if (threadIdx.x == 0) atomicExch(dev_error, 199); return; }
  // M-tiles are the fast grid dimension: the CTAs resident at any time share a few B (weight) tiles and sweep A,
  // which is the smaller operand and fits L2 -> each weight tile is fetched from HBM once.
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * TC_BN;
  const int total_kb = (a.Kd + TC_BK - 1) / TC_BK;
  const int nsplit = a.splits > 1 ? a.splits : 1;
  const int kb_chunk = (total_kb + nsplit - 1) / nsplit;
  const int zsplit = blockIdx.z + a.split_z0;
  const int kb0 = zsplit * kb_chunk;
  int num_kb = total_kb - kb0; if (num_kb > kb_chunk) num_kb = kb_chunk; if (num_kb < 0) num_kb = 0;

  if (warp == 0 && lane_id() == 0) {
    tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
    if (Cfg::DUAL) { tma_prefetch_desc(&tmB2_hi); tma_prefetch_desc(&tmB2_lo); }
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane_id() == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % Cfg::STAGES;
        const uint32_t ph = (kb / Cfg::STAGES) & 1;
        if (!mbar_wait(&empty[s], ph ^ 1)) { atomicExch(dev_error, 101); break; }
        uint8_t* st = smem + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
        const int kc = (kb0 + kb) * TC_BK;
        tma_load_2d(st + 0 * TC_TILE_BYTES, &tmA_hi, &full[s], kc, m0);
        tma_load_2d(st + 1 * TC_TILE_BYTES, &tmA_lo, &full[s], kc, m0);
        tma_load_2d(st + 2 * TC_TILE_BYTES, &tmB_hi, &full[s], kc, n0);
        tma_load_2d(st + 2 * TC_TILE_BYTES + Cfg::B_TILE_BYTES, &tmB_lo, &full[s], kc, n0);
        if (Cfg::DUAL) {
          tma_load_2d(st + 2 * TC_TILE_BYTES + 2 * Cfg::B_TILE_BYTES, &tmB2_hi, &full[s], kc, n0);
          tma_load_2d(st + 2 * TC_TILE_BYTES + 3 * Cfg::B_TILE_BYTES, &tmB2_lo, &full[s], kc, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the warp runs the loop convergently (uniform datapath), one elected lane issues =====
    {
      const uint32_t idesc = umma_idesc_tf32(TC_BM, TC_BN);
      const bool leader = elect_one();
      bool ok = true;
      for (int kb = 0; kb < num_kb && ok; ++kb) {
        const int s = kb % Cfg::STAGES;
        const uint32_t ph = (kb / Cfg::STAGES) & 1;
        // (no tcgen05.fence after this wait: the barrier is completed by TMA bytes, and a fence per k-block drains the
        //  MMA queue; all descriptors of a stage are its base descriptor plus compile-time constants)
        if (!mbar_wait(&full[s], ph)) { atomicExch(dev_error, 102); ok = false; break; }
        const uint64_t d0 = umma_desc_k128(smem_u32(smem + s * Cfg::STAGE_BYTES));
#pragma unroll
        for (int ks = 0; ks < TC_BK / 8; ++ks) {
          constexpr uint32_t T16 = TC_TILE_BYTES >> 4;   // descriptor address field counts 16-byte units
          constexpr uint32_t B16 = Cfg::B_TILE_BYTES >> 4;
          const uint32_t koff = ks * 2;                 // 8 tf32 = 32 bytes inside the 128-byte swizzle row
          const uint64_t a_hi = d0 + (uint64_t)(0 * T16 + koff);
          const uint64_t a_lo = d0 + (uint64_t)(1 * T16 + koff);
          const uint64_t b_hi = d0 + (uint64_t)(2 * T16 + koff);
          const uint64_t b_lo = d0 + (uint64_t)(2 * T16 + B16 + koff);
          const bool first = (kb < Cfg::NACC && ks == 0);          // first MMA into this accumulator overwrites
          const uint32_t acc_t = tmem_base + (uint32_t)(kb % Cfg::NACC) * Cfg::ACC_COLS;
          if (leader) {
            umma_tf32(acc_t, a_lo, b_hi, idesc, !first);
            umma_tf32(acc_t, a_hi, b_lo, idesc, true);
            umma_tf32(acc_t, a_hi, b_hi, idesc, true);
          }
          if (Cfg::DUAL) {
            const uint64_t c_hi = d0 + (uint64_t)(2 * T16 + 2 * B16 + koff);
            const uint64_t c_lo = d0 + (uint64_t)(2 * T16 + 3 * B16 + koff);
            if (leader) {
              umma_tf32(acc_t + TC_BN, a_lo, c_hi, idesc, !first);
              umma_tf32(acc_t + TC_BN, a_hi, c_lo, idesc, true);
              umma_tf32(acc_t + TC_BN, a_hi, c_hi, idesc, true);
            }
          }
        }
        if (leader) tc_commit(&empty[s]);     // frees the smem stage when these MMAs retire
        __syncwarp();
      }
      if (leader) tc_commit(acc_full);        // accumulator complete
      __syncwarp();
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int m = m0 + q * 32 + lane_id();
    float dsum = 0.f;
    bool ok = mbar_wait(acc_full, 0);
    if (!ok && lane_id() == 0) atomicExch(dev_error, 103);
    tc_fence_after();
    if (ok) {
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < TC_BN; c += 16) {
        float v[16], v2[16];
        tmem_ld16(trow + c, v);
        if (Cfg::DUAL) tmem_ld16(trow + TC_BN + c, v2);
        tc_wait_ld();
        const int nacc = num_kb < Cfg::NACC ? num_kb : Cfg::NACC;   // accumulators that received at least one k-block
        for (int q2 = 1; q2 < nacc; ++q2) {
          float t1[16], t2[16];
          tmem_ld16(trow + q2 * Cfg::ACC_COLS + c, t1);
          if (Cfg::DUAL) tmem_ld16(trow + q2 * Cfg::ACC_COLS + TC_BN + c, t2);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) { v[i] += t1[i]; if (Cfg::DUAL) v2[i] += t2[i]; }
        }
        if (num_kb == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) { v[i] = 0.f; v2[i] = 0.f; }
        }
        const int n = n0 + c;
        if (m < a.M_valid) {
          if (EPI == EPI_STORE) {
            if (n < a.N_valid) {
              float4* dst = reinterpret_cast<float4*>(a.C + (size_t)zsplit * a.split_stride + (size_t)m * a.ldc + n);
              if (a.bias) {
                const float4* bb = reinterpret_cast<const float4*>(a.bias + n);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 t4 = __ldg(bb + i);
                  v[4 * i] += t4.x; v[4 * i + 1] += t4.y; v[4 * i + 2] += t4.z; v[4 * i + 3] += t4.w;
                }
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
          } else if (EPI == EPI_GRAM) {
            if (n < a.N_valid) {
              float hi[16], lo[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) { hi[i] = epi_gram_value(m, n + i, a.R_valid, v[i]); lo[i] = tf32_lo(hi[i]); }
              float4* d1 = reinterpret_cast<float4*>(a.C + (size_t)m * a.ldc + n);
              float4* d2 = reinterpret_cast<float4*>(a.C_lo + (size_t)m * a.ldc + n);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                d1[i] = make_float4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                d2[i] = make_float4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
              }
            }
          } else if (EPI == EPI_MU_H) {
            // fused multiplicative update of H (rows = atoms m, columns = frames): 64 contiguous bytes of H per thread,
            // transposed copy with lanes = consecutive atoms
            if (n < a.N_valid) {
              float hi[16], lo[16];
              const float4* hp = reinterpret_cast<const float4*>(a.C + (size_t)m * a.ldc + n);
#pragma unroll
              for (int i = 0; i < 4; ++i) { const float4 t4 = hp[i]; hi[4 * i] = t4.x; hi[4 * i + 1] = t4.y; hi[4 * i + 2] = t4.z; hi[4 * i + 3] = t4.w; }
              const bool upd = !a.row_update || a.row_update[m];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const bool in = (n + i < a.N_valid);
                if (in && upd) hi[i] = hi[i] * v2[i] / fmaxf(v[i] + a.mu, a.flr);
                if (!in) hi[i] = 0.f;
                lo[i] = tf32_lo(hi[i]);
                dsum += hi[i];
              }
              float4* d1 = reinterpret_cast<float4*>(a.C + (size_t)m * a.ldc + n);
              float4* d2 = reinterpret_cast<float4*>(a.C_lo + (size_t)m * a.ldc + n);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                d1[i] = make_float4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                d2[i] = make_float4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
              }
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (n + i < a.N_valid) { a.CT[(size_t)(n + i) * a.ldct + m] = hi[i]; a.CT_lo[(size_t)(n + i) * a.ldct + m] = lo[i]; }
            }
          } else if (EPI == EPI_RECON) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n + i < a.N_valid) {
                a.C[(size_t)m * a.ldc + n + i] = epi_irm_value(v[i], v2[i], a.square);
                if (a.C_S) { a.C_S[(size_t)m * a.ldc + n + i] = v[i]; a.C_N[(size_t)m * a.ldc + n + i] = v2[i]; }
              }
          } else {
            // sparse-NMF reconstruction: Lambda = max(W H, flr) in both layouts (+ tf32 remainders) and the squared
            // error against V.  Row-major: 64 contiguous bytes per thread; transposed: lanes = consecutive rows.
            if (n < a.N_valid) {
              float hi[16], lo[16], vr[16];
              const float4* vp = reinterpret_cast<const float4*>(a.Vref + (size_t)m * a.ldv + n);
#pragma unroll
              for (int i = 0; i < 4; ++i) { const float4 t4 = __ldg(vp + i); vr[4 * i] = t4.x; vr[4 * i + 1] = t4.y; vr[4 * i + 2] = t4.z; vr[4 * i + 3] = t4.w; }
              float qv[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const bool in = (n + i < a.N_valid);
                hi[i] = in ? fmaxf(v[i], a.flr) : 0.f;
                if (EPI == EPI_LAMBDA_B) {         // beta != 2: P = L^(beta-1) replaces Lambda, Q = V L^(beta-2) replaces V
                  float P = 0.f, Q = 0.f, d = 0.f;
                  if (in) epi_beta_values(hi[i], vr[i], a.beta, P, Q, d);
                  hi[i] = P; qv[i] = Q; dsum += d;
                } else if (in) { const float d = vr[i] - hi[i]; dsum = fmaf(d, d, dsum); }
                lo[i] = tf32_lo(hi[i]);
              }
              float4* d1 = reinterpret_cast<float4*>(a.C + (size_t)m * a.ldc + n);
              float4* d2 = reinterpret_cast<float4*>(a.C_lo + (size_t)m * a.ldc + n);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                d1[i] = make_float4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                d2[i] = make_float4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
              }
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (n + i < a.N_valid) { a.CT[(size_t)(n + i) * a.ldct + m] = hi[i]; a.CT_lo[(size_t)(n + i) * a.ldct + m] = lo[i]; }
              if (EPI == EPI_LAMBDA_B) {
                float4* q1 = reinterpret_cast<float4*>(a.Q + (size_t)m * a.ldc + n);
                float4* q2 = reinterpret_cast<float4*>(a.Q_lo + (size_t)m * a.ldc + n);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  q1[i] = make_float4(qv[4 * i], qv[4 * i + 1], qv[4 * i + 2], qv[4 * i + 3]);
                  q2[i] = make_float4(tf32_lo(qv[4 * i]), tf32_lo(qv[4 * i + 1]), tf32_lo(qv[4 * i + 2]), tf32_lo(qv[4 * i + 3]));
                }
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (n + i < a.N_valid) { a.QT[(size_t)(n + i) * a.ldct + m] = qv[i]; a.QT_lo[(size_t)(n + i) * a.ldct + m] = tf32_lo(qv[i]); }
              }
            }
          }
        }
      }
    }
    if (EPI == EPI_LAMBDA || EPI == EPI_LAMBDA_B || EPI == EPI_MU_H) {
      float* red = reinterpret_cast<float*>(tmem_slot + 2);
      for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
      if (lane_id() == 0) red[q] = dsum;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 2 && lane_id() == 0)
        a.div_partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = (double)red[0] + (double)red[1] + (double)red[2] + (double)red[3];
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<Cfg::TMEM_COLS>(tmem_base); }
}

template <int EPI, int TC_BN>
static int launch_tc_impl(const GemmArgs& a, cudaStream_t st, int* dev_error) {
  using Cfg = TcCfg<EPI, TC_BN>;
  CUtensorMap tA_hi, tA_lo, tB_hi, tB_lo, tB2_hi, tB2_lo;
  int rc;
  if ((rc = make_tmap_2d(&tA_hi, a.A_hi, a.Kd, a.M, a.lda, TC_BK, TC_BM))) return rc;
  if ((rc = make_tmap_2d(&tA_lo, a.A_lo, a.Kd, a.M, a.lda, TC_BK, TC_BM))) return rc;
  if ((rc = make_tmap_2d(&tB_hi, a.B_hi, a.Kd, a.N, a.ldb, TC_BK, TC_BN))) return rc;
  if ((rc = make_tmap_2d(&tB_lo, a.B_lo, a.Kd, a.N, a.ldb, TC_BK, TC_BN))) return rc;
  tB2_hi = tB_hi; tB2_lo = tB_lo;
  if (Cfg::DUAL) {
    if ((rc = make_tmap_2d(&tB2_hi, a.B2_hi, a.Kd, a.N, a.ldb, TC_BK, TC_BN))) return rc;
    if ((rc = make_tmap_2d(&tB2_lo, a.B2_lo, a.Kd, a.N, a.ldb, TC_BK, TC_BN))) return rc;
  }
  // function attributes are per context: keep one flag per device
  static bool attr_set[DRNMF_MAX_DEVICES] = {};
  int dev = 0;
  DRNMF_CUDA(cudaGetDevice(&dev));
  DRNMF_CHECK(dev >= 0 && dev < DRNMF_MAX_DEVICES, "device ordinal %d out of range", dev);
  if (!attr_set[dev]) {
    DRNMF_CUDA(cudaFuncSetAttribute(k_gemm_tc<EPI, TC_BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[dev] = true;
  }
  dim3 grid((a.M + TC_BM - 1) / TC_BM, (a.N + TC_BN - 1) / TC_BN, a.split_nz > 0 ? a.split_nz : (a.splits > 1 ? a.splits : 1));
  k_gemm_tc<EPI, TC_BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(tA_hi, tA_lo, tB_hi, tB_lo, tB2_hi, tB2_lo, a, dev_error);
  count_launch();
  DRNMF_CUDA(cudaGetLastError());
  return DRNMF_OK;
}

// lazily allocated error word for handle-less GEMM users, one per device (device memory is per device)
static int* g_gemm_dev_error[DRNMF_MAX_DEVICES] = {};

static int gemm_error_word(int** out) {
  int dev = 0;
  DRNMF_CUDA(cudaGetDevice(&dev));
  DRNMF_CHECK(dev >= 0 && dev < DRNMF_MAX_DEVICES, "device ordinal %d out of range", dev);
  if (!g_gemm_dev_error[dev]) {
    DRNMF_CUDA(cudaMalloc(&g_gemm_dev_error[dev], sizeof(int)));
    DRNMF_CUDA(cudaMemset(g_gemm_dev_error[dev], 0, sizeof(int)));
  }
  *out = g_gemm_dev_error[dev];
  return DRNMF_OK;
}

int launch_gemm_tc(GemmEpi epi, const GemmArgs& a, cudaStream_t st) {
  DRNMF_CHECK(a.lda % 4 == 0 && a.ldb % 4 == 0, "tcgen05 GEMM needs row strides that are multiples of 4 floats");
  if (epi == EPI_STORE || epi == EPI_GRAM) DRNMF_CHECK(a.ldc % 4 == 0 && a.N_valid % 16 == 0, "tcgen05 GEMM store needs ldc%%4==0, N%%16==0");
  if (epi == EPI_MU_H) DRNMF_CHECK(a.ldc % 4 == 0, "EPI_MU_H needs ldc%%4==0");
  if (epi == EPI_LAMBDA || epi == EPI_LAMBDA_B) DRNMF_CHECK(a.ldc % 4 == 0 && a.ldv % 4 == 0, "EPI_LAMBDA needs ldc%%4==0 and ldv%%4==0");
  int* errw = nullptr;
  int rc = gemm_error_word(&errw);
  if (rc) return rc;
  // wide tiles when the grid still fills the device (and never for the dual-B epilogue)
  static const int force_bn = getenv("DRNMF_GEMM_BN") ? atoi(getenv("DRNMF_GEMM_BN")) : 0;
  const int m_plan = a.M_plan > 0 ? a.M_plan : a.M;        // a row block of a larger product keeps that product's tile shape
  const long long ctas256 = (long long)((m_plan + TC_BM - 1) / TC_BM) * ((a.N + 255) / 256) * (a.splits > 1 ? a.splits : 1);
  const bool wide = force_bn ? force_bn == 256 : (a.N >= 256 && ctas256 >= 4 * 148);     // >= 4 waves: no tail effect (measured:
                                                                                           // split-K weight-gradient GEMMs of 256 CTAs lose)
  switch (epi) {
    case EPI_STORE: return wide ? launch_tc_impl<EPI_STORE, 256>(a, st, errw) : launch_tc_impl<EPI_STORE, 128>(a, st, errw);
    case EPI_GRAM:  return wide ? launch_tc_impl<EPI_GRAM, 256>(a, st, errw) : launch_tc_impl<EPI_GRAM, 128>(a, st, errw);
    case EPI_RECON: return launch_tc_impl<EPI_RECON, 128>(a, st, errw);
    case EPI_MU_H: return launch_tc_impl<EPI_MU_H, 128>(a, st, errw);
    case EPI_LAMBDA: return wide ? launch_tc_impl<EPI_LAMBDA, 256>(a, st, errw) : launch_tc_impl<EPI_LAMBDA, 128>(a, st, errw);
    case EPI_LAMBDA_B: return wide ? launch_tc_impl<EPI_LAMBDA_B, 256>(a, st, errw) : launch_tc_impl<EPI_LAMBDA_B, 128>(a, st, errw);
  }
  return DRNMF_ERR_INVALID;
}

// Reads the current device's GEMM error word and, when it is set, clears it again (on the stream): a transient
// watchdog expiry is reported once and does not poison every later call of the process.
int gemm_device_error(cudaStream_t st) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= DRNMF_MAX_DEVICES) return -1;
  int* w = g_gemm_dev_error[dev];
  if (!w) return 0;
  int v = 0;
  if (cudaMemcpyAsync(&v, w, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
  if (v != 0) cudaMemsetAsync(w, 0, sizeof(int), st);
  return v;
}

}  // namespace drnmf
