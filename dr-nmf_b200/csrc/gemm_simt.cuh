// CUDA-core fp32 tile main loop: C[m][n] = sum_k A[m][k] * B[n][k]  (both operands K-major).
// This is the "semantics lock" path (DRNMF_IMPL_SIMT): exact fp32 FMAs, same data layouts as the tcgen05 path.
#pragma once
#include "common.cuh"

namespace drnmf {

constexpr int SIMT_BM = 64, SIMT_BN = 64, SIMT_BK = 16, SIMT_THREADS = 256;

// 256 threads; thread (ty = tid/16, tx = tid%16) owns rows m0+ty*4+i, cols n0+tx*4+j.
template <bool DUAL>
__device__ __forceinline__ void simt_tile_mainloop(const float* __restrict__ A, int lda, int M,
                                                   const float* __restrict__ B, const float* __restrict__ B2, int ldb,
                                                   int N, int Kd, int m0, int n0, float (&acc)[4][4],
                                                   float (&acc2)[4][4]) {
  __shared__ float As[SIMT_BK][SIMT_BM + 4];
  __shared__ float Bs[SIMT_BK][SIMT_BN + 4];
  __shared__ float Bs2[DUAL ? SIMT_BK : 1][SIMT_BN + 4];
  const int tid = threadIdx.x;
  const int lr = tid >> 2, lk = (tid & 3) * 4;   // loader: row 0..63, k offset 0,4,8,12
  const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; acc2[i][j] = 0.f; }
  for (int k0 = 0; k0 < Kd; k0 += SIMT_BK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, b2 = a;
    if (m0 + lr < M && k0 + lk < Kd) a = *reinterpret_cast<const float4*>(A + (size_t)(m0 + lr) * lda + k0 + lk);
    if (n0 + lr < N && k0 + lk < Kd) {
      b = *reinterpret_cast<const float4*>(B + (size_t)(n0 + lr) * ldb + k0 + lk);
      if (DUAL) b2 = *reinterpret_cast<const float4*>(B2 + (size_t)(n0 + lr) * ldb + k0 + lk);
    }
    As[lk + 0][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
    Bs[lk + 0][lr] = b.x; Bs[lk + 1][lr] = b.y; Bs[lk + 2][lr] = b.z; Bs[lk + 3][lr] = b.w;
    if (DUAL) { Bs2[lk + 0][lr] = b2.x; Bs2[lk + 1][lr] = b2.y; Bs2[lk + 2][lr] = b2.z; Bs2[lk + 3][lr] = b2.w; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SIMT_BK; ++kk) {
      float av[4], bv[4], bv2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) { bv[j] = Bs[kk][tx * 4 + j]; if (DUAL) bv2[j] = Bs2[kk][tx * 4 + j]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
          if (DUAL) acc2[i][j] = fmaf(av[i], bv2[j], acc2[i][j]);
        }
    }
    __syncthreads();
  }
}

// ---- element epilogues shared by the SIMT and tcgen05 GEMMs ---------------------------------------
// Gram: S_k^T[j][i] = delta_ij - acc, zero outside the valid R x R block (enhance.py:172-181).
__device__ __forceinline__ float epi_gram_value(int m, int n, int R_valid, float acc) {
  if (m >= R_valid || n >= R_valid) return 0.f;
  return ((m == n) ? 1.f : 0.f) - acc;
}
// DivideAbyAplusB (custom_layers.py:41-45): exp(log(1e-7 + A) - log(1e-7 + A + B)), optional 'square' transform (mode 1).
// mode 2: the SNMF baseline's ratio mask  A / (1e-9 + A + B)  (enhance.py:848-852).
// beta-divergence pieces of one element (sparse_nmf_gpu.m:212-276, beta != 2): lam = max(W H, flr), v = V (made
// positive on entry, :201-205).  P = lam^(beta-1), Q = v lam^(beta-2), d = the element's divergence.
__device__ __forceinline__ void epi_beta_values(float lam, float v, float beta, float& P, float& Q, float& d) {
  if (beta == 1.f) {            // KL (:212-216, :269)
    const float r = v / lam;
    P = 1.f; Q = r; d = v * logf(r) - v + lam;
  } else if (beta == 0.f) {     // IS (:222-226 with beta = 0, :273)
    const float r = v / lam;
    P = 1.f / lam; Q = r / lam; d = r - logf(r) - 1.f;
  } else {                      // generic (:222-226, :275-276)
    P = powf(lam, beta - 1.f); Q = v * powf(lam, beta - 2.f);
    d = (powf(v, beta) + (beta - 1.f) * powf(lam, beta) - beta * v * P) / (beta * (beta - 1.f));
  }
}

__device__ __forceinline__ float epi_irm_value(float s, float n, int mode) {
  if (mode == 2) return s / (1e-9f + s + n);
  if (mode == 1) { s *= s; n *= n; }
  return expf(logf(1e-7f + s) - logf(1e-7f + s + n));
}

}  // namespace drnmf
