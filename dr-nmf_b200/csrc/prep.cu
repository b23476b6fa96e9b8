// Parameter maps of build_alt (enhance.py:139-206) and the Masking layer (enhance.py:253) as HBM-bound
// elementwise / transpose kernels.  All outputs are written in the K-major, zero-padded layouts the tensor-core
// GEMMs consume, together with their tf32 remainders (the "lo" operand of the 3xTF32 product).
#include "internal.h"

namespace drnmf {

// inv_norm[k][j] = 1 / sqrt(sum_f exp(log_D[k][f][j])^2)          (enhance.py:175: K.sqrt(K.sum(K.square(...), axis=0)))
// grid (ceil(R/32), nD), block (32, 8): coalesced along j, 8-way split over f.
__global__ void k_colnorm(const float* __restrict__ log_D, int F, int R, int Rp, float* __restrict__ inv_norm) {
  __shared__ float red[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  const float* base = log_D + (size_t)blockIdx.y * F * R;
  float s = 0.f;
  if (j < R) {
    for (int f = threadIdx.y; f < F; f += 8) {
      float e = expf(base[(size_t)f * R + j]);
      s = fmaf(e, e, s);
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
    if (j < Rp) inv_norm[(size_t)blockIdx.y * Rp + j] = (j < R) ? 1.0f / sqrtf(t) : 0.f;
  }
}

// Dt[k][j][f] = exp(log_D[kD][f][j]) * inv_norm ;  Wt[k][j][f] = Dt / alph[k][j]      (32x32 smem transpose)
// grid (Fp/32, Rp/32, K), block (32, 8)
__global__ void k_prep_dict(const float* __restrict__ log_D, int n_log_D, const float* __restrict__ log_alph,
                            int n_log_alph, int alph_dim, const float* __restrict__ inv_norm, int F, int R, int Rp,
                            int Fp, float* __restrict__ Dt_hi, float* __restrict__ Dt_lo, float* __restrict__ Wt_hi,
                            float* __restrict__ Wt_lo, float* __restrict__ Dm_hi, float* __restrict__ Dm_lo,
                            float* __restrict__ alph_out) {
  __shared__ float tile[32][33];
  const int k = blockIdx.z;
  const int kD = (n_log_D == 1) ? 0 : k;
  const int kA = (n_log_alph == 1) ? 0 : k;
  const int f0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  const float* src = log_D + (size_t)kD * F * R;
  for (int y = threadIdx.y; y < 32; y += 8) {
    int f = f0 + y, j = j0 + threadIdx.x;
    const float e = (f < F && j < R) ? expf(src[(size_t)f * R + j]) : 0.f;
    tile[y][threadIdx.x] = e;
    const float dn = (j < R) ? e * inv_norm[(size_t)kD * Rp + j] : 0.f;       // D^_k[f][j], rows f (backward operand)
    Dm_hi[((size_t)k * Fp + f) * Rp + j] = dn; Dm_lo[((size_t)k * Fp + f) * Rp + j] = tf32_lo(dn);
    if (f == 0) alph_out[(size_t)k * Rp + j] = (j < R) ? expf(log_alph[(size_t)kA * alph_dim + (alph_dim == 1 ? 0 : j)]) : 1.f;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    int j = j0 + y, f = f0 + threadIdx.x;
    float d = 0.f, w = 0.f;
    if (j < R && f < F) {
      d = tile[threadIdx.x][y] * inv_norm[(size_t)kD * Rp + j];
      float la = log_alph[(size_t)kA * alph_dim + (alph_dim == 1 ? 0 : j)];
      w = d / expf(la);
    }
    size_t o = ((size_t)k * Rp + j) * Fp + f;
    Dt_hi[o] = d; Dt_lo[o] = tf32_lo(d);
    Wt_hi[o] = w; Wt_lo[o] = tf32_lo(w);
  }
}

// b_k[j] = -exp(log_lam1_k)/exp(log_alph_kj) ; h0 = softplus(log_h0)
__global__ void k_prep_bias_h0(const float* __restrict__ log_alph, int n_log_alph, int alph_dim,
                               const float* __restrict__ log_lam1, int n_log_lam1, const float* __restrict__ log_h0,
                               int R, int Rp, int K, float* __restrict__ bias, float* __restrict__ h0) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < K * Rp) {
    int k = idx / Rp, j = idx % Rp;
    float v = 0.f;
    if (j < R) {
      float la = log_alph[(size_t)(n_log_alph == 1 ? 0 : k) * alph_dim + (alph_dim == 1 ? 0 : j)];
      float ll = log_lam1[n_log_lam1 == 1 ? 0 : k];
      v = -expf(ll) / expf(la);
    }
    bias[idx] = v;
  }
  if (idx < Rp) {
    float v = 0.f;
    if (idx < R) {
      float x = log_h0[idx];
      v = (x > 20.f) ? x : log1pf(expf(x));
    }
    h0[idx] = v;
  }
}

// EcT[f][j] = exp(k_clean[j][f]) (j < r) ; EnT[f][j] = exp(k_noise[j-r][f]) (r <= j < R) ; zero elsewhere.
// grid (Rp/32, Fq/32), block (32, 8)
__global__ void k_prep_recon(const float* __restrict__ k_clean, const float* __restrict__ k_noise, int F, int R, int r,
                             int Rp, int Fp, float* __restrict__ EcT_hi, float* __restrict__ EcT_lo,
                             float* __restrict__ EnT_hi, float* __restrict__ EnT_lo, float* __restrict__ EcB_hi,
                             float* __restrict__ EcB_lo) {
  __shared__ float tc[32][33], tn[32][33];
  const int j0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  for (int y = threadIdx.y; y < 32; y += 8) {
    int j = j0 + y, f = f0 + threadIdx.x;
    float c = 0.f, n = 0.f;
    if (f < F) {
      if (j < r) c = expf(k_clean[(size_t)j * F + f]);
      else if (j < R) n = expf(k_noise[(size_t)(j - r) * F + f]);
    }
    tc[y][threadIdx.x] = c; tn[y][threadIdx.x] = n;
    if (f < Fp) {   // block-diagonal copy, rows j: [clean | 0] for j < r, [0 | noise] for r <= j < R
      const size_t o = (size_t)j * (2 * Fp);
      EcB_hi[o + f] = c; EcB_lo[o + f] = tf32_lo(c);
      EcB_hi[o + Fp + f] = n; EcB_lo[o + Fp + f] = tf32_lo(n);
    }
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    int f = f0 + y, j = j0 + threadIdx.x;
    float c = tc[threadIdx.x][y], n = tn[threadIdx.x][y];
    size_t o = (size_t)f * Rp + j;
    EcT_hi[o] = c; EcT_lo[o] = tf32_lo(c);
    EnT_hi[o] = n; EnT_lo[o] = tf32_lo(n);
  }
}

int launch_prep_params(drnmf_handle* h, const float* log_D, int n_log_D, const float* log_alph, int n_log_alph,
                       int alph_dim, const float* log_lam1, int n_log_lam1, const float* log_h0, const float* k_clean,
                       const float* k_noise, cudaStream_t st) {
  const int F = h->F, R = h->R, K = h->K, Rp = h->Rp, Fp = h->Fp;
  k_colnorm<<<dim3((Rp + 31) / 32, n_log_D), dim3(32, 8), 0, st>>>(log_D, F, R, Rp, h->inv_norm);
  k_prep_dict<<<dim3(Fp / 32, Rp / 32, K), dim3(32, 8), 0, st>>>(log_D, n_log_D, log_alph, n_log_alph, alph_dim,
                                                                 h->inv_norm, F, R, Rp, Fp, h->Dt_hi, h->Dt_lo,
                                                                 h->Wt_hi, h->Wt_lo, h->Dm_hi, h->Dm_lo, h->alph);
  k_prep_bias_h0<<<(K * Rp + 255) / 256, 256, 0, st>>>(log_alph, n_log_alph, alph_dim, log_lam1, n_log_lam1, log_h0, R,
                                                       Rp, K, h->bias, h->h0);
  k_prep_recon<<<dim3(Rp / 32, h->Fq / 32), dim3(32, 8), 0, st>>>(k_clean, k_noise, F, R, h->r, Rp, Fp, h->EcT_hi,
                                                                  h->EcT_lo, h->EnT_hi, h->EnT_lo, h->EcB_hi, h->EcB_lo);
  DRNMF_CUDA(cudaMemsetAsync(h->log_h0, 0, sizeof(float) * Rp, st));
  DRNMF_CUDA(cudaMemcpyAsync(h->log_h0, log_h0, sizeof(float) * R, cudaMemcpyDeviceToDevice, st));
  h->n_log_D = n_log_D; h->n_log_alph = n_log_alph; h->alph_dim = alph_dim; h->n_log_lam1 = n_log_lam1;
  count_launch(4);
  DRNMF_CUDA(cudaGetLastError());
  return DRNMF_OK;
}

// Keras Masking(mask_value) (enhance.py:253): m = any_f(x != mask_value), x~ = x * m.  One warp per frame row;
// writes the zero-padded row (Fp wide) and its tf32 remainder.
__global__ void k_mask_pad(const float* __restrict__ x, int BT, int F, int Fp, float mask_value,
                           float* __restrict__ xp_hi, float* __restrict__ xp_lo, float* __restrict__ mvalid, int B_tm,
                           int t_begin, int t_count) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  size_t orow = row;                                   // output row: b-major like the input, or time-major (t*B + b)
  if (B_tm > 0) {                                      // frames [t_begin, t_begin + t_count) of every utterance
    if (row >= B_tm * t_count) return;
    const int T = BT / B_tm, b = row / t_count, t = t_begin + (row - b * t_count);
    row = b * T + t; orow = (size_t)t * B_tm + b;
  } else if (row >= BT) return;
  const int lane = threadIdx.x & 31;
  const float* src = x + (size_t)row * F;
  bool any = false;
  for (int f = lane; f < F; f += 32) any |= (src[f] != mask_value);
  any = __any_sync(0xffffffffu, any);
  for (int f = lane; f < Fp; f += 32) {
    float v = (f < F && any) ? src[f] : 0.f;
    xp_hi[orow * Fp + f] = v;
    xp_lo[orow * Fp + f] = tf32_lo(v);
  }
  if (lane == 0) mvalid[row] = any ? 1.f : 0.f;
}

int launch_mask_pad(const drnmf_handle* h, const float* x, int BT, float mask_value, FwdWorkspace& w, cudaStream_t st,
                    int B_tmajor, int t_begin, int t_count) {
  const int rows = B_tmajor > 0 ? B_tmajor * (t_count > 0 ? t_count : BT / B_tmajor) : BT;
  if (B_tmajor > 0 && t_count <= 0) { t_begin = 0; t_count = BT / B_tmajor; }
  k_mask_pad<<<(rows + 7) / 8, 256, 0, st>>>(x, BT, h->F, h->Fp, mask_value, w.xp_hi, w.xp_lo, w.mvalid, B_tmajor, t_begin, t_count);
  count_launch();
  DRNMF_CUDA(cudaGetLastError());
  return DRNMF_OK;
}

}  // namespace drnmf
