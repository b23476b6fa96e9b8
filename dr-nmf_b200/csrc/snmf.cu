// "Well-done" sparse NMF, Euclidean branch (sparseNMF/sparse_nmf_gpu.m:163-298, beta = 2) as fused GEMM + elementwise
// kernels.  Per iteration:
//   H <- H .* (W^T V) ./ max(W^T L + mu, flr)          (:217-221)      two (n x F).(F x R) GEMMs + k_mu_h
//   L <- max(W H, flr)                                  (:228)          GEMM with the EPI_LAMBDA epilogue
//   W <- W .* (V H^T + W .* colsum(L H^T .* W)) ./ max(L H^T + W .* colsum(V H^T .* W), flr)   (:243-249)
//                                                                       two split-K (F x n).(n x R) GEMMs + k_mu_w
//   W <- W ./ colnorm(W) ; L <- max(W H, flr)           (:262-263)
//   div = |V - L|_F^2 (no 1/2) ; cost = div + mu * sum(H)   (:271,278) ; stop when |d cost| / cost_prev < conv_eps (:288-296)
// beta != 2 (KL, IS, generic beta-divergence; :212-216, :222-226, :232-241, :250-259, :266-276): the same iteration with
// P = L^(beta-1) in the place of L and Q = V .* L^(beta-2) in the place of V - the lambda GEMM's epilogue
// (EPI_LAMBDA_B) writes P and Q in all layouts and sums the divergence, everything else is shared.  (For beta = 1 the
// .m file writes W^T 1 and 1 H^T as column / row sums; P = 1 gives the same numbers through the GEMM.)
// Both operand layouts of every matrix are kept (frame-major for the contractions over F and R, feature-major for the
// contractions over frames) together with their tf32 remainders; the epilogues write all of them.
#include "internal.h"

#include <cmath>
#include <cstdio>

namespace drnmf {

struct SnmfWs {
  // K-major operands (hi, lo).  Fk = ceil32(F), Rk = ceil32(R), nk = ceil128(n)
  float *Wm_hi, *Wm_lo;   // F x Rk      W          (B operand of L = H W^T)
  float *WT_hi, *WT_lo;   // R x Fk      W^T        (B operand of W^T L, W^T V)
  float *Ht_hi, *Ht_lo;   // nk x Rk     H^T        (A operand of L)
  float *Hm_hi, *Hm_lo;   // R x nk      H          (B operand of V H^T, L H^T)
  float *Vt_hi, *Vt_lo;   // nk x Fk     V^T
  float *Vm_hi, *Vm_lo;   // F x nk      V
  float *Lt_hi, *Lt_lo;   // nk x Fk     L^T
  float *Lm_hi, *Lm_lo;   // F x nk      L
  float *Qt_hi, *Qt_lo;   // nk x Fk     (V .* L^(beta-2))^T    beta != 2 only (L* then hold P = L^(beta-1))
  float *Qm_hi, *Qm_lo;   // F x nk      V .* L^(beta-2)
  float* vmin;            // smallest positive entry of V (:201-205)
  float *dph, *dmh;       // nk x Rk     (W^T L)^T , (W^T V)^T
  float *VHp, *LHp;       // splits x F x Rk   split-K partials of V H^T, L H^T
  double *div_part, *hsum_part, *scal;   // per-CTA partial sums; scal[0] = div, scal[1] = mu * sum(H)
  float* wnorm;           // Rk
  size_t bytes;
  int Fk, Rk, nk, splits, n_div_part, n_h_part;
};

static size_t al256(size_t x) { return (x + 255) / 256 * 256; }

static SnmfWs carve_snmf(int F, int n, int R, void* base, float beta = 2.f) {
  SnmfWs w;
  w.Fk = round_up(F, 32); w.Rk = round_up(R, 32); w.nk = round_up(n, 128);
  int total_kb = w.nk / 32;
  // split-K: parallelism for the 5 x 8 output tiles AND short tensor-core accumulation chains (<= 64 k-blocks per split,
  // dealt to 4 accumulators: the round-toward-zero bias of tcgen05 accumulation grows with the chain length)
  w.splits = total_kb >= 64 ? (total_kb + 63) / 64 : (total_kb >= 8 ? 4 : 1);
  if (w.splits < 16 && total_kb >= 64) w.splits = 16;
  if (w.splits > 128) w.splits = 128;
  uint8_t* p = (uint8_t*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* q = p ? p + off : nullptr; off += al256(bytes); return q; };
  const size_t F_ = F, R_ = R, Fk = w.Fk, Rk = w.Rk, nk = w.nk;
  w.Wm_hi = (float*)take(F_ * Rk * 4); w.Wm_lo = (float*)take(F_ * Rk * 4);
  w.WT_hi = (float*)take(R_ * Fk * 4); w.WT_lo = (float*)take(R_ * Fk * 4);
  w.Ht_hi = (float*)take(nk * Rk * 4); w.Ht_lo = (float*)take(nk * Rk * 4);
  w.Hm_hi = (float*)take(R_ * nk * 4); w.Hm_lo = (float*)take(R_ * nk * 4);
  w.Vt_hi = (float*)take(nk * Fk * 4); w.Vt_lo = (float*)take(nk * Fk * 4);
  w.Vm_hi = (float*)take(F_ * nk * 4); w.Vm_lo = (float*)take(F_ * nk * 4);
  w.Lt_hi = (float*)take(nk * Fk * 4); w.Lt_lo = (float*)take(nk * Fk * 4);
  w.Lm_hi = (float*)take(F_ * nk * 4); w.Lm_lo = (float*)take(F_ * nk * 4);
  w.Qt_hi = w.Qt_lo = w.Qm_hi = w.Qm_lo = nullptr;
  if (beta != 2.f) {
    w.Qt_hi = (float*)take(nk * Fk * 4); w.Qt_lo = (float*)take(nk * Fk * 4);
    w.Qm_hi = (float*)take(F_ * nk * 4); w.Qm_lo = (float*)take(F_ * nk * 4);
  }
  w.vmin = (float*)take(256);
  w.dph = (float*)take(nk * Rk * 4); w.dmh = (float*)take(nk * Rk * 4);
  w.VHp = (float*)take((size_t)w.splits * F_ * Rk * 4); w.LHp = (float*)take((size_t)w.splits * F_ * Rk * 4);
  w.n_div_part = (int)((nk / 64) * ((Fk + 63) / 64 + 1));            // upper bound over both GEMM tilings
  w.n_h_part = (int)((nk + 31) / 32 * ((Rk + 31) / 32));
  w.div_part = (double*)take((size_t)w.n_div_part * 8);
  w.hsum_part = (double*)take((size_t)w.n_h_part * 8);
  w.scal = (double*)take(64);
  w.wnorm = (float*)take(Rk * 4);
  w.bytes = off;
  return w;
}

// ---- layout kernels (32x32 smem-tile transposes, coalesced on both sides) -------------------------------
// src (rows x cols, ld_src) -> dst_rm (rows x ld_rm, zero padded), dst_tr (cols x ld_tr, zero padded), with remainders
__global__ void k_split_both(const float* __restrict__ src, int rows, int cols, int ld_src, float* __restrict__ rm_hi,
                             float* __restrict__ rm_lo, int rm_rows, int ld_rm, float* __restrict__ tr_hi,
                             float* __restrict__ tr_lo, int tr_rows, int ld_tr, const float* __restrict__ row_scale,
                             const float* __restrict__ col_scale, int scale_is_div,
                             const float* __restrict__ zero_fill = nullptr) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int r = r0 + y, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = src[(size_t)r * ld_src + c];
      if (row_scale) v = scale_is_div ? v / row_scale[r] : v * row_scale[r];
      if (col_scale) v = scale_is_div ? v / col_scale[c] : v * col_scale[c];
      if (zero_fill && v == 0.f) v = *zero_fill;          // sparse_nmf_gpu.m:201-205
    }
    tile[y][threadIdx.x] = v;
    if (rm_hi && r < rm_rows && c < ld_rm) { rm_hi[(size_t)r * ld_rm + c] = v; rm_lo[(size_t)r * ld_rm + c] = tf32_lo(v); }
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int c = c0 + y, r = r0 + threadIdx.x;
    if (tr_hi && c < tr_rows && r < ld_tr) {
      const float v = tile[threadIdx.x][y];
      tr_hi[(size_t)c * ld_tr + r] = v; tr_lo[(size_t)c * ld_tr + r] = tf32_lo(v);
    }
  }
}

// smallest positive entry (positive floats order like their bit patterns): *out starts at +inf
__global__ void k_min_positive(const float* __restrict__ src, size_t n, unsigned* __restrict__ out) {
  unsigned m = 0x7f800000u;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = src[i];
    if (v > 0.f) m = min(m, __float_as_uint(v));
  }
  for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m != 0x7f800000u) atomicMin(out, m);
}

// column l2 norms of a row-major (rows x cols) matrix: grid ceil(cols/32), block (32, 8)
__global__ void k_colnorm_rm(const float* __restrict__ src, int rows, int cols, int ld, float* __restrict__ norm) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < cols)
    for (int r = threadIdx.y; r < rows; r += 8) { const float v = src[(size_t)r * ld + c]; s = fmaf(v, v, s); }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
    for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
    norm[c] = sqrtf(t);
  }
}

// H update on the frame-major layout: Ht[n][r] *= dmh / max(dph + mu, flr) for updated rows r; writes both layouts and
// remainders, and the per-block partial of sum(H) (all entries, as in cost = div + sum(sparsity .* h)).
// grid (Rk/32, nk/32), block (32, 8)
__global__ void k_mu_h(float* __restrict__ Ht_hi, float* __restrict__ Ht_lo, float* __restrict__ Hm_hi,
                       float* __restrict__ Hm_lo, const float* __restrict__ dph, const float* __restrict__ dmh, int n,
                       int R, int Rk, int nk, float mu, float flr, const uint8_t* __restrict__ h_update, int do_update,
                       double* __restrict__ hsum_part) {
  __shared__ float tile[32][33];
  __shared__ float red[8];
  const int r0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  float local = 0.f;
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int fr = n0 + y, r = r0 + threadIdx.x;
    float h = 0.f;
    if (fr < n && r < R) {
      const size_t o = (size_t)fr * Rk + r;
      h = Ht_hi[o];
      if (do_update && (!h_update || h_update[r])) h = h * dmh[o] / fmaxf(dph[o] + mu, flr);
      Ht_hi[o] = h; Ht_lo[o] = tf32_lo(h);
      local += h;
    }
    tile[y][threadIdx.x] = h;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int r = r0 + y, fr = n0 + threadIdx.x;
    if (r < R && fr < nk) {
      const float h = tile[threadIdx.x][y];
      Hm_hi[(size_t)r * nk + fr] = h; Hm_lo[(size_t)r * nk + fr] = tf32_lo(h);
    }
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if (threadIdx.x == 0) red[threadIdx.y] = local;
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    double t = 0.0;
    for (int y = 0; y < 8; ++y) t += (double)red[y];
    hsum_part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

// W update for a block of 32 columns (all F rows): reduce split-K partials in split order, column sums, multiplicative
// update of the updated columns, renormalisation of every column (:262), both layouts + remainders.
// grid ceil(R/32), block (32, ny <= 32).  W is read from / written to Wm (F x Rk).
__global__ void __launch_bounds__(1024) k_mu_w(float* __restrict__ Wm_hi, float* __restrict__ Wm_lo, float* __restrict__ WT_hi,
                       float* __restrict__ WT_lo, float* __restrict__ VHp, float* __restrict__ LHp, int splits, int F, int R,
                       int Rk, int Fk, float flr, const uint8_t* __restrict__ w_update) {
  __shared__ float red_v[32][33], red_l[32][33];
  __shared__ float sv[32], sl[32], nrm[32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int ny = blockDim.y;                                 // rows are dealt to the ny thread rows (<= 32)
  const size_t sstride = (size_t)F * Rk;
  const bool upd = (c < R) && (!w_update || w_update[c]);
  // pass 1: VH, LH = sum over splits (stored back into split 0), column sums of VH.*W and LH.*W
  float a_v = 0.f, a_l = 0.f;
  if (c < R)
    for (int f = threadIdx.y; f < F; f += ny) {
      const size_t o = (size_t)f * Rk + c;
      float vh = 0.f, lh = 0.f;
      for (int s = 0; s < splits; ++s) { vh += VHp[s * sstride + o]; lh += LHp[s * sstride + o]; }
      VHp[o] = vh; LHp[o] = lh;
      const float w = Wm_hi[o];
      a_v = fmaf(vh, w, a_v); a_l = fmaf(lh, w, a_l);
    }
  red_v[threadIdx.y][threadIdx.x] = a_v; red_l[threadIdx.y][threadIdx.x] = a_l;
  __syncthreads();
  if (threadIdx.y == 0) {
    float tv = 0.f, tl = 0.f;
    for (int y = 0; y < ny; ++y) { tv += red_v[y][threadIdx.x]; tl += red_l[y][threadIdx.x]; }
    sv[threadIdx.x] = tv; sl[threadIdx.x] = tl;
  }
  __syncthreads();
  // pass 2: multiplicative update, accumulate the new column norm
  float a_n = 0.f;
  if (c < R)
    for (int f = threadIdx.y; f < F; f += ny) {
      const size_t o = (size_t)f * Rk + c;
      float w = Wm_hi[o];
      if (upd) {
        const float dpw = fmaxf(LHp[o] + sv[threadIdx.x] * w, flr);
        const float dmw = VHp[o] + sl[threadIdx.x] * w;
        w = w * dmw / dpw;
        Wm_hi[o] = w;
      }
      a_n = fmaf(w, w, a_n);
    }
  red_v[threadIdx.y][threadIdx.x] = a_n;
  __syncthreads();
  if (threadIdx.y == 0) {
    float t = 0.f;
    for (int y = 0; y < ny; ++y) t += red_v[y][threadIdx.x];
    nrm[threadIdx.x] = sqrtf(t);
  }
  __syncthreads();
  // pass 3: normalise, write both layouts and remainders
  if (c < Rk)
    for (int f = threadIdx.y; f < F; f += ny) {
      const size_t o = (size_t)f * Rk + c;
      const float w = (c < R) ? Wm_hi[o] / nrm[threadIdx.x] : 0.f;
      Wm_hi[o] = w; Wm_lo[o] = tf32_lo(w);
      if (c < R) { WT_hi[(size_t)c * Fk + f] = w; WT_lo[(size_t)c * Fk + f] = tf32_lo(w); }
    }
}

// scal[0] = sum(div partials), scal[1] = mu * sum(hsum partials)   (single block, fixed order)
__global__ void k_mu_cost(const double* __restrict__ div_part, int n_div, const double* __restrict__ hsum_part, int n_h,
                          double mu, double* __restrict__ scal) {
  __shared__ double red[256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n_div; i += 256) a += div_part[i];
  for (int i = threadIdx.x; i < n_h; i += 256) b += hsum_part[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) scal[0] = red[0];
  __syncthreads();
  red[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) scal[1] = mu * red[0];
}

// copy back: W (F x R) from Wm, H (R x n) from Hm
__global__ void k_unpad(const float* __restrict__ src, int rows, int cols, int ld_src, float* __restrict__ dst) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (size_t)rows * cols) dst[idx] = src[(idx / cols) * ld_src + idx % cols];
}

static int run_gemm_impl(bool simt, GemmEpi epi, const GemmArgs& a, cudaStream_t st) {
  return simt ? launch_gemm_simt(epi, a, st) : launch_gemm_tc(epi, a, st);
}

size_t snmf_workspace_bytes(int F, int n, int R, float beta) { return carve_snmf(F, n, R, nullptr, beta).bytes; }

// sum the split-K partials into split 0 (fixed order), so that one buffer per matrix can be all-reduced across ranks
__global__ void k_reduce_splits(float* __restrict__ part, int splits, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = 0.f;
  for (int s = 0; s < splits; ++s) v += part[(size_t)s * n + i];
  part[i] = v;
}

int snmf_mu_ed(int F, int n, int R, float beta, const float* V, float* W, float* H, const uint8_t* w_update, const uint8_t* h_update,
               int any_w_update, int any_h_update, float sparsity, int max_iter, float conv_eps, double* cost_host,
               double* div_host, int* iters_host, void* ws, size_t ws_bytes, bool simt, cudaStream_t st,
               drnmf_allreduce_fn allreduce, void* user) {
  SnmfWs w = carve_snmf(F, n, R, ws, beta);
  const bool ed = (beta == 2.f);
  if (ws_bytes < w.bytes) { set_error("snmf workspace too small: need %zu bytes, got %zu", w.bytes, ws_bytes); return DRNMF_ERR_WORKSPACE; }
  const int Fk = w.Fk, Rk = w.Rk, nk = w.nk;
  const float flr = 1e-9f;
  const dim3 tb(32, 8);
  // every padded buffer starts from zero (padding rows / columns are operands of the contractions)
  DRNMF_CUDA(cudaMemsetAsync(ws, 0, w.bytes, st));
  // ---- init (:163-173): normalise the columns of W, rescale the rows of H, L = max(W H, flr) ----
  k_colnorm_rm<<<(R + 31) / 32, tb, 0, st>>>(W, F, R, R, w.wnorm);
  k_split_both<<<dim3((R + 31) / 32, (F + 31) / 32), tb, 0, st>>>(W, F, R, R, w.Wm_hi, w.Wm_lo, F, Rk, w.WT_hi, w.WT_lo, R, Fk,
                                                                  nullptr, w.wnorm, 1);
  k_split_both<<<dim3((n + 31) / 32, (R + 31) / 32), tb, 0, st>>>(H, R, n, n, w.Hm_hi, w.Hm_lo, R, nk, w.Ht_hi, w.Ht_lo, nk, Rk,
                                                                  w.wnorm, nullptr, 0);
  if (!ed) {   // :201-205  v(v==0) = min(v(v>0)); with sharded frames the minimum is over all ranks (dtype 2 = float32 MIN)
    const unsigned inf_bits = 0x7f800000u;
    DRNMF_CUDA(cudaMemcpyAsync(w.vmin, &inf_bits, 4, cudaMemcpyHostToDevice, st));
    k_min_positive<<<148 * 4, 256, 0, st>>>(V, (size_t)F * n, reinterpret_cast<unsigned*>(w.vmin));
    count_launch();
    if (allreduce && allreduce(user, w.vmin, 1, 2, st)) { set_error("all-reduce callback failed"); return DRNMF_ERR_CUDA; }
  }
  k_split_both<<<dim3((n + 31) / 32, (F + 31) / 32), tb, 0, st>>>(V, F, n, n, w.Vm_hi, w.Vm_lo, F, nk, w.Vt_hi, w.Vt_lo, nk, Fk,
                                                                  nullptr, nullptr, 0, ed ? nullptr : w.vmin);
  count_launch(4);
  DRNMF_CUDA(cudaGetLastError());

  // L = max(W-rows . H^T-rows, flr): (F x R).(n x R)^T, all layouts + divergence.  The bins are the M dimension: F = 513
  // costs 5 tiles of 128 rows (640) - as the N dimension of 256-column tiles it cost 3 tiles (768), a third of them for
  // the one Nyquist bin (2.35 -> 2.0 ms per launch at 513 x 225,000).
  auto lambda_gemm = [&]() {
    GemmArgs a{};
    a.A_hi = w.Wm_hi; a.A_lo = w.Wm_lo; a.lda = Rk;
    a.B_hi = w.Ht_hi; a.B_lo = w.Ht_lo; a.ldb = Rk;
    a.M = F; a.N = n; a.Kd = Rk; a.M_valid = F; a.N_valid = n;
    a.C = w.Lm_hi; a.C_lo = w.Lm_lo; a.ldc = nk; a.CT = w.Lt_hi; a.CT_lo = w.Lt_lo; a.ldct = Fk;
    a.Vref = w.Vm_hi; a.ldv = nk; a.div_partials = w.div_part; a.flr = flr;
    a.beta = beta; a.Q = w.Qm_hi; a.Q_lo = w.Qm_lo; a.QT = w.Qt_hi; a.QT_lo = w.Qt_lo;
    return run_gemm_impl(simt, ed ? EPI_LAMBDA : EPI_LAMBDA_B, a, st);
  };
  auto proj_gemm = [&](const float* A_hi, const float* A_lo, float* out) {   // (n x F).(R x F)^T -> n x Rk
    GemmArgs a{};
    a.A_hi = A_hi; a.A_lo = A_lo; a.lda = Fk;
    a.B_hi = w.WT_hi; a.B_lo = w.WT_lo; a.ldb = Fk;
    a.M = n; a.N = R; a.Kd = Fk; a.M_valid = n; a.N_valid = Rk;
    a.C = out; a.ldc = Rk;
    return run_gemm_impl(simt, EPI_STORE, a, st);
  };
  auto corr_gemm = [&](const float* A_hi, const float* A_lo, float* out) {   // (F x n).(R x n)^T -> splits x F x Rk
    GemmArgs a{};
    a.A_hi = A_hi; a.A_lo = A_lo; a.lda = nk;
    a.B_hi = w.Hm_hi; a.B_lo = w.Hm_lo; a.ldb = nk;
    a.M = F; a.N = R; a.Kd = nk; a.M_valid = F; a.N_valid = Rk;
    a.C = out; a.ldc = Rk; a.splits = w.splits; a.split_stride = (size_t)F * Rk;
    return run_gemm_impl(simt, EPI_STORE, a, st);
  };
  // H update fused into ONE dual-operand GEMM (both W^T P and W^T Q against the same W^T tiles, atoms = M dimension): the
  // epilogue applies h <- h .* (W^T Q) ./ max(W^T P + mu, flr) to the H tile, writes both layouts + remainders and the
  // tile's sum of H.  Saves the (n x R) round trip of the two projections and the separate pass of k_mu_h over H
  // (3.9 -> 3.1 ms of an 11.7 ms iteration at 513 x 225,000).  Used when W moves too (W^T Q is not constant).
  auto h_update_gemm = [&](const float* P_hi, const float* P_lo, const float* Q_hi, const float* Q_lo) {
    GemmArgs a{};
    a.A_hi = w.WT_hi; a.A_lo = w.WT_lo; a.lda = Fk;
    a.B_hi = P_hi; a.B_lo = P_lo; a.B2_hi = Q_hi; a.B2_lo = Q_lo; a.ldb = Fk;
    a.M = R; a.N = n; a.Kd = Fk; a.M_valid = R; a.N_valid = n;
    a.C = w.Hm_hi; a.C_lo = w.Hm_lo; a.ldc = nk; a.CT = w.Ht_hi; a.CT_lo = w.Ht_lo; a.ldct = Rk;
    a.div_partials = w.hsum_part; a.flr = flr; a.mu = sparsity; a.row_update = h_update;
    return run_gemm_impl(simt, EPI_MU_H, a, st);
  };
  const bool fused_h = !getenv("DRNMF_MU_UNFUSED");
  // number of div partials the lambda GEMM writes depends on the tiling of the implementation in use
  const int tile = simt ? 64 : 128;
  const int n_div = ((n + tile - 1) / tile) * ((F + tile - 1) / tile);
  const dim3 gh(Rk / 32, nk / 32);
  int n_h = gh.x * gh.y;                                           // partial sums of H written per iteration

  int rc;
  if ((rc = lambda_gemm())) return rc;
  // the operands that stand where V stands in the Euclidean updates
  const float *Xt_hi = ed ? w.Vt_hi : w.Qt_hi, *Xt_lo = ed ? w.Vt_lo : w.Qt_lo;
  const float *Xm_hi = ed ? w.Vm_hi : w.Qm_hi, *Xm_lo = ed ? w.Vm_lo : w.Qm_lo;
  const bool dmh_constant = ed && !any_w_update && any_h_update;                                       // W^T V is constant
  if (dmh_constant) { if ((rc = proj_gemm(w.Vt_hi, w.Vt_lo, w.dmh))) return rc; }
  double last_cost = INFINITY;
  int it = 0;
  for (it = 1; it <= max_iter; ++it) {
    if (any_h_update && !dmh_constant && fused_h) {
      if ((rc = h_update_gemm(w.Lt_hi, w.Lt_lo, Xt_hi, Xt_lo))) return rc;
      n_h = ((R + tile - 1) / tile) * ((n + tile - 1) / tile);
      if ((rc = lambda_gemm())) return rc;
    } else if (any_h_update) {
      if ((rc = proj_gemm(w.Lt_hi, w.Lt_lo, w.dph))) return rc;
      if (!dmh_constant) { if ((rc = proj_gemm(Xt_hi, Xt_lo, w.dmh))) return rc; }
      k_mu_h<<<gh, tb, 0, st>>>(w.Ht_hi, w.Ht_lo, w.Hm_hi, w.Hm_lo, w.dph, w.dmh, n, R, Rk, nk, sparsity, flr, h_update, 1, w.hsum_part);
      count_launch();
      if ((rc = lambda_gemm())) return rc;
    } else if (it == 1) {   // sum(H) is constant: compute it once with a no-op update mask
      k_mu_h<<<gh, tb, 0, st>>>(w.Ht_hi, w.Ht_lo, w.Hm_hi, w.Hm_lo, w.dph, w.dmh, n, R, Rk, nk, sparsity, flr, h_update, 0, w.hsum_part);
      count_launch();
    }
    if (any_w_update) {
      if ((rc = corr_gemm(Xm_hi, Xm_lo, w.VHp))) return rc;
      if ((rc = corr_gemm(w.Lm_hi, w.Lm_lo, w.LHp))) return rc;
      // split-K partials -> split 0 with one thread per element (fixed split order): the update kernel below has only
      // ceil(R/32) CTAs, and summing the partials there made it the largest item of an iteration (830 us of 2.2 ms)
      const size_t ne = (size_t)F * Rk;
      if (w.splits > 1) {
        k_reduce_splits<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(w.VHp, w.splits, ne);
        k_reduce_splits<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(w.LHp, w.splits, ne);
        count_launch(2);
      }
      if (allreduce) {   // frames are sharded over ranks: V H^T and L H^T are sums over ALL frames
        if (allreduce(user, w.VHp, ne, 0, st) || allreduce(user, w.LHp, ne, 0, st)) { set_error("all-reduce callback failed"); return DRNMF_ERR_CUDA; }
      }
      k_mu_w<<<(Rk + 31) / 32, dim3(32, 32), 0, st>>>(w.Wm_hi, w.Wm_lo, w.WT_hi, w.WT_lo, w.VHp, w.LHp, 1, F, R, Rk, Fk, flr, w_update);
      count_launch();
      if ((rc = lambda_gemm())) return rc;
    }
    k_mu_cost<<<1, 256, 0, st>>>(w.div_part, n_div, w.hsum_part, n_h, (double)sparsity, w.scal);
    count_launch();
    if (allreduce && allreduce(user, w.scal, 2, 1, st)) { set_error("all-reduce callback failed"); return DRNMF_ERR_CUDA; }
    double sc[2];
    DRNMF_CUDA(cudaMemcpyAsync(sc, w.scal, sizeof(sc), cudaMemcpyDeviceToHost, st));
    DRNMF_CUDA(cudaStreamSynchronize(st));
    const double div = sc[0], cost = sc[0] + sc[1];
    div_host[it - 1] = div; cost_host[it - 1] = cost;
    if (it > 1 && conv_eps > 0) {
      const double e = fabs(cost - last_cost) / last_cost;
      if (e < conv_eps) { ++it; break; }
    }
    last_cost = cost;
  }
  *iters_host = it - 1;
  k_unpad<<<(unsigned)(((size_t)F * R + 255) / 256), 256, 0, st>>>(w.Wm_hi, F, R, Rk, W);
  k_unpad<<<(unsigned)(((size_t)R * n + 255) / 256), 256, 0, st>>>(w.Hm_hi, R, n, nk, H);
  count_launch(2);
  DRNMF_CUDA(cudaGetLastError());
  return DRNMF_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Frame-parallel ISTA with a tied dictionary (enhance.py:402-418, `ista_ed`: dead code in the reference, oracle "A"):
//   xhat = W H ; repeat K times:  H <- max(0, -lam1/alph + H + (1/alph) W^T (x - xhat)) ; xhat = W H
// Frames are independent, so this is two GEMMs + two elementwise kernels per iteration on the frame-major layout.
__global__ void k_ista_residual(const float* __restrict__ xt, const float* __restrict__ xhat, size_t n, float* __restrict__ e_hi,
                                float* __restrict__ e_lo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float e = xt[i] - xhat[i]; e_hi[i] = e; e_lo[i] = tf32_lo(e); }
}
__global__ void k_ista_update(float* __restrict__ Ht_hi, float* __restrict__ Ht_lo, const float* __restrict__ G, int n, int R,
                              int Rk, float lam_over_alph, float inv_alph) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * Rk) return;
  const int r = (int)(i % Rk);
  float h = 0.f;
  if (r < R) h = fmaxf(0.f, -lam_over_alph + Ht_hi[i] + inv_alph * G[i]);
  Ht_hi[i] = h; Ht_lo[i] = tf32_lo(h);
}

size_t ista_workspace_bytes(int F, int n, int R) {
  const size_t Fk = round_up(F, 32), Rk = round_up(R, 32), nk = round_up(n, 128);
  return al256((size_t)F * Rk * 4) * 2 + al256((size_t)R * Fk * 4) * 2 + al256(nk * Rk * 4) * 3 + al256(nk * Fk * 4) * 4 + 4096;
}

int ista_ed(int F, int n, int R, const float* x, const float* W, float* H, float lam1, float alph, int iters, void* ws,
            size_t ws_bytes, bool simt, cudaStream_t st) {
  const int Fk = round_up(F, 32), Rk = round_up(R, 32), nk = round_up(n, 128);
  if (ws_bytes < ista_workspace_bytes(F, n, R)) { set_error("ista workspace too small"); return DRNMF_ERR_WORKSPACE; }
  uint8_t* p = (uint8_t*)ws;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* q = p + off; off += al256(bytes); return (float*)q; };
  float *Wm_hi = take((size_t)F * Rk * 4), *Wm_lo = take((size_t)F * Rk * 4);
  float *WT_hi = take((size_t)R * Fk * 4), *WT_lo = take((size_t)R * Fk * 4);
  float *Ht_hi = take((size_t)nk * Rk * 4), *Ht_lo = take((size_t)nk * Rk * 4), *G = take((size_t)nk * Rk * 4);
  float *xt = take((size_t)nk * Fk * 4), *xhat = take((size_t)nk * Fk * 4), *e_hi = take((size_t)nk * Fk * 4), *e_lo = take((size_t)nk * Fk * 4);
  DRNMF_CUDA(cudaMemsetAsync(ws, 0, off, st));
  const dim3 tb(32, 8);
  k_split_both<<<dim3((R + 31) / 32, (F + 31) / 32), tb, 0, st>>>(W, F, R, R, Wm_hi, Wm_lo, F, Rk, WT_hi, WT_lo, R, Fk, nullptr, nullptr, 0);
  k_split_both<<<dim3((n + 31) / 32, (R + 31) / 32), tb, 0, st>>>(H, R, n, n, nullptr, nullptr, 0, 0, Ht_hi, Ht_lo, nk, Rk, nullptr, nullptr, 0);
  k_split_both<<<dim3((n + 31) / 32, (F + 31) / 32), tb, 0, st>>>(x, F, n, n, nullptr, nullptr, 0, 0, xt, e_lo, nk, Fk, nullptr, nullptr, 0);
  count_launch(3);
  int rc;
  for (int it = 0; it < iters; ++it) {
    GemmArgs a{};      // xhat^T = H^T-rows . W-rows
    a.A_hi = Ht_hi; a.A_lo = Ht_lo; a.lda = Rk; a.B_hi = Wm_hi; a.B_lo = Wm_lo; a.ldb = Rk;
    a.M = n; a.N = F; a.Kd = Rk; a.M_valid = n; a.N_valid = Fk; a.C = xhat; a.ldc = Fk;
    if ((rc = run_gemm_impl(simt, EPI_STORE, a, st))) return rc;
    k_ista_residual<<<(unsigned)(((size_t)nk * Fk + 255) / 256), 256, 0, st>>>(xt, xhat, (size_t)nk * Fk, e_hi, e_lo);
    GemmArgs g{};      // G^T = (x - xhat)^T-rows . W^T-rows
    g.A_hi = e_hi; g.A_lo = e_lo; g.lda = Fk; g.B_hi = WT_hi; g.B_lo = WT_lo; g.ldb = Fk;
    g.M = n; g.N = R; g.Kd = Fk; g.M_valid = n; g.N_valid = Rk; g.C = G; g.ldc = Rk;
    if ((rc = run_gemm_impl(simt, EPI_STORE, g, st))) return rc;
    k_ista_update<<<(unsigned)(((size_t)n * Rk + 255) / 256), 256, 0, st>>>(Ht_hi, Ht_lo, G, n, R, Rk, lam1 / alph, 1.f / alph);
    count_launch(2);
  }
  // H (R x n) <- Ht (n x Rk): transposed copy through the tile kernel (hi only; the lo output goes to scratch)
  k_split_both<<<dim3((Rk + 31) / 32, (n + 31) / 32), tb, 0, st>>>(Ht_hi, n, R, Rk, nullptr, nullptr, 0, 0, H, G, R, n, nullptr, nullptr, 0);
  count_launch();
  DRNMF_CUDA(cudaGetLastError());
  return DRNMF_OK;
}

// ---- SNMF ratio mask (enhance.py:847-852): irm = S^/(1e-9 + S^ + N^), S^ = W[:, :r] H[:r], N^ = W[:, r:] H[r:] -------------
// One dual-B GEMM with the ratio epilogue: A = W (F x Rp, K-major over the atoms), B = H_clean^T, B2 = H_noise^T
// (n x Rp, the other source's columns zeroed), C = irm (F x n).
// H (R x n) -> HcT / HnT (np x Rp) hi | lo: 32 x 32 tiled transpose
__global__ void k_irm_prep_h(const float* __restrict__ H, int R, int r, int n, int Rp, float* __restrict__ HcT_hi,
                             float* __restrict__ HcT_lo, float* __restrict__ HnT_hi, float* __restrict__ HnT_lo) {
  __shared__ float tile[32][33];
  const int j0 = blockIdx.y * 32, c0 = blockIdx.x * 32;       // atoms, frames
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int j = j0 + y, c = c0 + threadIdx.x;
    tile[y][threadIdx.x] = (j < R && c < n) ? H[(size_t)j * n + c] : 0.f;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int c = c0 + y, j = j0 + threadIdx.x;
    if (j >= Rp) continue;
    const float v = tile[threadIdx.x][y];
    const float vc = (j < r) ? v : 0.f, vn = (j >= r && j < R) ? v : 0.f;
    const size_t o = (size_t)c * Rp + j;
    HcT_hi[o] = vc; HcT_lo[o] = tf32_lo(vc); HnT_hi[o] = vn; HnT_lo[o] = tf32_lo(vn);
  }
}
__global__ void k_irm_prep_w(const float* __restrict__ W, int F, int R, int Rp, float* __restrict__ A_hi, float* __restrict__ A_lo) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)F * Rp) return;
  const int f = (int)(idx / Rp), j = (int)(idx % Rp);
  const float v = j < R ? W[(size_t)f * R + j] : 0.f;
  A_hi[idx] = v; A_lo[idx] = tf32_lo(v);
}

size_t snmf_irm_workspace_bytes(int F, int n, int R) {
  const size_t Rp = round_up(R, 32), np_ = round_up(n, 128);
  return 4 * round_up_sz(np_ * Rp * 4, 256) + 2 * round_up_sz((size_t)F * Rp * 4, 256);
}

int snmf_irm(int F, int n, int R, int r, const float* W, const float* H, float* irm, void* ws, size_t ws_bytes, bool simt,
             cudaStream_t st) {
  const size_t need = snmf_irm_workspace_bytes(F, n, R);
  if (ws_bytes < need) { set_error("snmf_irm workspace too small: need %zu bytes, got %zu", need, ws_bytes); return DRNMF_ERR_WORKSPACE; }
  const int Rp = round_up(R, 32);
  const size_t np_ = round_up(n, 128), hb = round_up_sz(np_ * Rp * 4, 256), wb = round_up_sz((size_t)F * Rp * 4, 256);
  uint8_t* p = (uint8_t*)ws;
  float *HcT_hi = (float*)p, *HcT_lo = (float*)(p + hb), *HnT_hi = (float*)(p + 2 * hb), *HnT_lo = (float*)(p + 3 * hb);
  float *A_hi = (float*)(p + 4 * hb), *A_lo = (float*)(p + 4 * hb + wb);
  k_irm_prep_h<<<dim3((unsigned)(np_ / 32), (Rp + 31) / 32), dim3(32, 8), 0, st>>>(H, R, r, n, Rp, HcT_hi, HcT_lo, HnT_hi, HnT_lo);
  k_irm_prep_w<<<(unsigned)(((size_t)F * Rp + 255) / 256), 256, 0, st>>>(W, F, R, Rp, A_hi, A_lo);
  count_launch(2);
  DRNMF_CUDA(cudaGetLastError());
  GemmArgs a{};
  a.A_hi = A_hi; a.A_lo = A_lo; a.lda = Rp;
  a.B_hi = HcT_hi; a.B_lo = HcT_lo; a.B2_hi = HnT_hi; a.B2_lo = HnT_lo; a.ldb = Rp;
  a.M = F; a.N = n; a.Kd = Rp; a.C = irm; a.ldc = n; a.M_valid = F; a.N_valid = n;
  a.square = 2;                       // ratio epilogue S / (1e-9 + S + N)
  return simt ? launch_gemm_simt(EPI_RECON, a, st) : launch_gemm_tc(EPI_RECON, a, st);
}

}  // namespace drnmf
