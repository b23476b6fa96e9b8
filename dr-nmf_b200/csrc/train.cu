// Training step of the unfolded network (enhance.py:1040-1073): forward with stored activations, masked-MSE loss,
// and the backward pass -- BPTT over the (T x K_layers) chain plus the recurrence-free weight-gradient contractions --
// down to gradients of the reference's own parameters (log_D_k, log_alph_k, log_lam1, log_h0, recon kernels).
// The reference has no hand-written backward (Theano autodiff through scan); the maths is derived in DESIGN.md.
//
// Loss (build-defined normalisation):  L_sum = sum_{b,t} m[b,t] * mean_f (x*irm - y)^2 ; the caller divides
// by sum m (all-reduced over ranks in data-parallel training).  All gradients below are d L_sum / d parameter.
//
// Frame order of every K = (frames) contraction is time-major (t*Bp + b): any order gives the same sum, and this one
// lets the forward/backward kernels write the transposed activations as float4 over the batch.
#include "internal.h"
#include "gemm_simt.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace drnmf {

static size_t al2(size_t x) { return (x + 255) / 256 * 256; }

struct TrainWs {
  FwdWorkspace fwd;
  float *S, *N, *irm;          // BT x F   raw reconstructions and mask
  float *dSN_hi, *dSN_lo;      // BTp x 2Fp   [dL/dS | dL/dN], frame-major (b*T + t)
  float *dSNT_hi, *dSNT_lo;    // 2Fp x TB    same, transposed, time-major frames
  float *dH;                   // BTp x Rp    dL/dH from the head
  float *xT_hi, *xT_lo;        // Fx x TB     masked input transposed (time-major) + a row of ones (bias gradient)
  float *deltaT_hi, *deltaT_lo;// K x Rp x TB dL/dz^k
  float *G, *dbuf;             // Bp x Rp ; 2 x Bp x Rp   state-gradient carry ; delta ping-pong (frame-major)
  float *psum2;                // 2 x 256 x 2 x Bp  partial row sums of the persistent backward chain
  float *rs_part;              // K x NC x Bp  partial row sums of delta^k (rank-1 leak gradient), NC = Rp/64 column tiles
  float *part;                 // split-K partials (max over the GEMMs that use them)
  float *partS_all, *partX_all;// K x splits x Rp x Rp and K x splits x Rp x Fx: per-layer partials of the weight-gradient
                               // GEMMs when their late-frame blocks are computed under the backward chain
  unsigned int* progress;      // frames the backward chain has completely processed (device word)
  float *gsym_hi, *gsym_lo;    // Rp x Rp
  float *dDt;                  // Rp x Fp
  float *dEc;                  // 2 x Rp x Fp   H^T dS , H^T dN
  double *loss_part, *scal;    // per-block loss partials ; [0] = L_sum, [1] = sum m
  float *rowacc;               // Rp x 4       per-row partials of the scalar gradients (alph terms, lam)
  float *rowS;                 // Rp x (Rp/32) per-row, per-column-tile partials of the S term of d log alph
  size_t bytes;
  int Fx, TB, splits_w, splits_x;
};

static TrainWs carve_train(const drnmf_handle* h, int B, int T, void* base) {
  TrainWs w;
  w.fwd = carve_forward_ws(h, B, T, base);
  uint8_t* p = (uint8_t*)base;
  size_t off = w.fwd.bytes;
  auto take = [&](size_t bytes) { void* q = p ? p + off : nullptr; off += al2(bytes); return q; };
  const size_t BT = (size_t)B * T, BTp = round_up_sz(BT, 128), F = h->F, Fp = h->Fp, Rp = h->Rp, K = h->K;
  const size_t Bp = w.fwd.Bp, TB = (size_t)T * Bp;
  w.TB = (int)TB;
  w.Fx = round_up(h->F + 1, 32);
  w.fwd.actT_hi = (float*)take(K * Rp * TB * 4);
  w.fwd.actT_lo = (float*)take(K * Rp * TB * 4);
  w.S = (float*)take(BT * F * 4); w.N = (float*)take(BT * F * 4); w.irm = (float*)take(BT * F * 4);
  w.dSN_hi = (float*)take(BTp * 2 * Fp * 4); w.dSN_lo = (float*)take(BTp * 2 * Fp * 4);
  w.dSNT_hi = (float*)take(2 * Fp * TB * 4); w.dSNT_lo = (float*)take(2 * Fp * TB * 4);
  w.dH = (float*)take(BTp * Rp * 4);
  w.xT_hi = (float*)take((size_t)w.Fx * TB * 4); w.xT_lo = (float*)take((size_t)w.Fx * TB * 4);
  w.deltaT_hi = (float*)take(K * Rp * TB * 4); w.deltaT_lo = (float*)take(K * Rp * TB * 4);
  w.G = (float*)take(Bp * Rp * 4); w.dbuf = (float*)take(2 * Bp * Rp * 4);
  w.psum2 = (float*)take(2 * 256 * 2 * Bp * 4);
  w.rs_part = (float*)take(K * (Rp / SIMT_BN) * Bp * 4);
  const int kb = (int)(TB / 32);
  w.splits_w = kb >= 256 ? 8 : (kb >= 64 ? 4 : (kb >= 16 ? 2 : 1));
  w.splits_x = w.splits_w;
  const size_t part_elems = (size_t)w.splits_w * Rp * (Rp > (size_t)w.Fx ? Rp : (size_t)w.Fx);
  w.part = (float*)take(part_elems * 4);
  w.partS_all = nullptr; w.partX_all = nullptr;
  if (w.splits_w >= 4) {      // the pipelined plan needs >= 4 split-K blocks (kb >= 64: every training-sized batch)
    w.partS_all = (float*)take((size_t)K * w.splits_w * Rp * Rp * 4);
    w.partX_all = (float*)take((size_t)K * w.splits_x * Rp * (size_t)w.Fx * 4);
  }
  w.progress = (unsigned int*)take(256);
  w.gsym_hi = (float*)take(Rp * Rp * 4); w.gsym_lo = (float*)take(Rp * Rp * 4);
  w.dDt = (float*)take(Rp * Fp * 4);
  w.dEc = (float*)take(2 * Rp * Fp * 4);
  w.loss_part = (double*)take(((BT + 7) / 8) * 16);
  w.scal = (double*)take(64);
  w.rowacc = (float*)take(Rp * 4 * 4);
  w.rowS = (float*)take(Rp * (Rp / 32) * 4);
  w.bytes = off;
  return w;
}

size_t train_workspace_bytes(const drnmf_handle* h, int B, int T) { return carve_train(h, B, T, nullptr).bytes; }

// ---- loss + head gradient --------------------------------------------------------------------------------------
// One warp per frame (b,t): err = x*irm - y ; L += m * mean_f err^2 ; dirm = m * 2 err x / F ;
// irm = A/Bq with A = eps + S', Bq = eps + S' + N' (S' = S or S^2): dS' = dirm * N'/Bq^2, dN' = -dirm * A/Bq^2.
// Writes [dS|dN] frame-major (hi, lo) and transposed time-major (hi, lo).
// kind 1 = SNMF pretraining cost (enhance.py:1024-1036): 0.5 * mean_f (S + N - x)^2 per frame (the l1 term on H is added
// by k_l1_head); dS = dN = (S + N - x) / F.
__global__ void k_loss_head(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ S,
                            const float* __restrict__ N, const float* __restrict__ mvalid, int B, int T, int Bp, int F,
                            int Fp, int square, int kind, float* __restrict__ dSN_hi, float* __restrict__ dSN_lo,
                            float* __restrict__ dSNT_hi, float* __restrict__ dSNT_lo, double* __restrict__ loss_part) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  __shared__ double red[8][2];
  double lsum = 0.0;
  float mv = 0.f;
  if (row < B * T) {
    const int b = row / T, t = row % T;
    mv = mvalid[row];
    const size_t TB = (size_t)T * Bp, col = (size_t)t * Bp + b;
    float acc = 0.f;
    for (int f = lane; f < Fp; f += 32) {
      float dS = 0.f, dN = 0.f;
      if (f < F && mv != 0.f) {
        const size_t o = (size_t)row * F + f;
        float s = S[o], n = N[o];
        const float s0 = s, n0 = n;
        if (kind == 1) {
          const float err = s + n - x[o];
          acc = fmaf(0.5f * err, err, acc);
          dS = err / (float)F; dN = dS;
        } else {
        if (square) { s *= s; n *= n; }
        const float A = 1e-7f + s, Bq = A + n;
        const float irm = expf(logf(A) - logf(Bq));
        const float xv = x[o], err = xv * irm - y[o];
        acc = fmaf(err, err, acc);
        const float dirm = 2.f * err * xv / (float)F;
        dS = dirm * n / (Bq * Bq);
        dN = -dirm * A / (Bq * Bq);
        if (square) { dS *= 2.f * s0; dN *= 2.f * n0; }
        }
      }
      const size_t o2 = (size_t)row * (2 * Fp);
      dSN_hi[o2 + f] = dS; dSN_lo[o2 + f] = tf32_lo(dS);
      dSN_hi[o2 + Fp + f] = dN; dSN_lo[o2 + Fp + f] = tf32_lo(dN);
      dSNT_hi[(size_t)f * TB + col] = dS; dSNT_lo[(size_t)f * TB + col] = tf32_lo(dS);
      dSNT_hi[(size_t)(Fp + f) * TB + col] = dN; dSNT_lo[(size_t)(Fp + f) * TB + col] = tf32_lo(dN);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    lsum = (double)acc / (double)F;
  }
  if (lane == 0) { red[threadIdx.x >> 5][0] = lsum; red[threadIdx.x >> 5][1] = (double)mv; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < 8; ++w) { a += red[w][0]; c += red[w][1]; }
    loss_part[2 * (size_t)blockIdx.x] = a; loss_part[2 * (size_t)blockIdx.x + 1] = c;
  }
}

// l1 term of the SNMF pretraining cost: per valid frame  lam1 * (R/F) * mean_j |H_j| = (lam1/F) * sum_j H_j  (H >= 0);
// dH[bt][j] += (lam1/F) [H_j > 0].  One warp per frame; per-block partial sums go to loss_part (pairs: value, 0).
__global__ void k_l1_head(const float* __restrict__ H, const float* __restrict__ mvalid, int BT, int R, int Rp, float coef,
                          float* __restrict__ dH, double* __restrict__ loss_part) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  __shared__ double red[8];
  float sum = 0.f;
  if (row < BT && mvalid[row] != 0.f) {
    for (int j = lane; j < R; j += 32) {
      const float h = H[(size_t)row * Rp + j];
      sum += h;
      if (h > 0.f) dH[(size_t)row * Rp + j] += coef;
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  }
  if (lane == 0) red[threadIdx.x >> 5] = (double)sum * coef;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < 8; ++w) a += red[w];
    loss_part[2 * (size_t)blockIdx.x] = a; loss_part[2 * (size_t)blockIdx.x + 1] = 0.0;
  }
}
__global__ void k_add_scal(double* __restrict__ scal) { scal[0] += scal[2]; }

__global__ void k_reduce_pairs(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double ra[256], rb[256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) { a += part[2 * i]; b += part[2 * i + 1]; }
  ra[threadIdx.x] = a; rb[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) { ra[threadIdx.x] += ra[threadIdx.x + o]; rb[threadIdx.x] += rb[threadIdx.x + o]; } __syncthreads(); }
  if (threadIdx.x == 0) { out[0] = ra[0]; out[1] = rb[0]; }
}

// masked input transposed to time-major frames, plus a row of ones at f == F (bias gradient): grid (TB/32, Fx/32)
__global__ void k_xT(const float* __restrict__ xp, int B, int T, int Bp, int F, int Fp, int Fx, float* __restrict__ xT_hi,
                     float* __restrict__ xT_lo, int xp_tmajor) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, f0 = blockIdx.y * 32;     // c = t*Bp + b
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int c = c0 + y, f = f0 + threadIdx.x;
    const int t = c / Bp, b = c % Bp;
    float v = 0.f;
    if (b < B && t < T) {
      if (f < F) v = xp[(xp_tmajor ? (size_t)t * B + b : (size_t)b * T + t) * Fp + f];
      else if (f == F) v = 1.f;
    }
    tile[y][threadIdx.x] = v;
  }
  __syncthreads();
  const size_t TB = (size_t)T * Bp;
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int f = f0 + y, c = c0 + threadIdx.x;
    if (f < Fx && (size_t)c < TB) { const float v = tile[threadIdx.x][y]; xT_hi[(size_t)f * TB + c] = v; xT_lo[(size_t)f * TB + c] = tf32_lo(v); }
  }
}

// ---- backward chain (CUDA-core steps; the chain has the same all-to-all structure as the forward recurrence) -------
struct BwdArgs {
  const float *ST, *actT, *mvalid, *dH;
  const float* alph;   // K x Rp step sizes when alph is a vector (untie_alph), else null: see k_bwd_step
  float *deltaT_hi, *deltaT_lo, *G, *dbuf, *rs_part;
  int B, Bp, T, K, R, Rp, t, k, nc;
  size_t TB;
  float d0mo, o0, ok;
};

// start of frame t: dg^{K-1} = m * (dH[b,t] + G[b]) ; delta^{K-1} = dg .* 1[act^{K-1} > 0].  One CTA per utterance.
__global__ void k_bwd_frame_begin(BwdArgs a) {
  const int b = blockIdx.x;
  const size_t bt = (size_t)b * a.T + a.t;
  const float mv = a.mvalid[bt];
  __shared__ float red[8];
  float local = 0.f;
  const size_t col = (size_t)a.t * a.Bp + b;
  for (int j = threadIdx.x; j < a.Rp; j += blockDim.x) {
    float d = 0.f;
    if (mv != 0.f && j < a.R) {
      const float act = a.actT[((size_t)(a.K - 1) * a.Rp + j) * a.TB + col];
      if (act > 0.f) d = a.dH[bt * a.Rp + j] + a.G[(size_t)b * a.Rp + j];
    }
    // slot 0: operand of the first product (pre-scaled by 1/alph for vector alph, see k_bwd_step)
    a.dbuf[(size_t)b * a.Rp + j] = (a.alph && a.K > 1) ? d / a.alph[(size_t)(a.K - 1) * a.Rp + j] : d;
    const size_t o = ((size_t)(a.K - 1) * a.Rp + j) * a.TB + col;
    a.deltaT_hi[o] = d; a.deltaT_lo[o] = tf32_lo(d);
    local += d;
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    a.rs_part[((size_t)(a.K - 1) * a.nc + 0) * a.Bp + b] = t;
    for (int c = 1; c < a.nc; ++c) a.rs_part[((size_t)(a.K - 1) * a.nc + c) * a.Bp + b] = 0.f;
  }
}

// layer k -> k-1:  dg^{k-1}[b][i] = sum_j delta^k[b][j] S_k[i][j] ;  delta^{k-1} = dg .* 1[act^{k-1} > 0]
// S_k is symmetric for scalar alph: the stored S_k^T serves both directions.  For vector alph (untie_alph)
// S_k^T = I - diag(1/alph) G with G symmetric, hence  sum_j delta_j S_k[i][j] = alph_i * sum_j (delta_j / alph_j) S_k^T[i][j]:
// the same stored matrix serves if the operand is pre-scaled by 1/alph_k and the product post-scaled by alph_k.
__global__ void __launch_bounds__(SIMT_THREADS) k_bwd_step(BwdArgs a) {
  float acc[4][4], acc2[4][4];
  const int m0 = blockIdx.y * SIMT_BM, n0 = blockIdx.x * SIMT_BN;      // m = utterance, n = input atom i
  const float* din = a.dbuf + (size_t)((a.K - 1 - a.k) & 1) * a.Bp * a.Rp;
  float* dout = a.dbuf + (size_t)((a.K - a.k) & 1) * a.Bp * a.Rp;
  simt_tile_mainloop<false>(din, a.Rp, a.B, a.ST, nullptr, a.Rp, a.Rp, a.Rp, m0, n0, acc, acc2);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  __shared__ float rs[SIMT_BM][17];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = m0 + ty * 4 + i;
    float part = 0.f;
    if (b < a.B) {
      const size_t col = (size_t)a.t * a.Bp + b;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = n0 + tx * 4 + jj;
        if (j >= a.Rp) continue;
        float d = 0.f;
        if (j < a.R) {
          const float act = a.actT[((size_t)(a.k - 1) * a.Rp + j) * a.TB + col];
          if (act > 0.f) d = a.alph ? acc[i][jj] * a.alph[(size_t)a.k * a.Rp + j] : acc[i][jj];
        }
        dout[(size_t)b * a.Rp + j] = (a.alph && a.k > 1) ? d / a.alph[(size_t)(a.k - 1) * a.Rp + j] : d;
        const size_t o = ((size_t)(a.k - 1) * a.Rp + j) * a.TB + col;
        a.deltaT_hi[o] = d; a.deltaT_lo[o] = tf32_lo(d);
        part += d;
      }
    }
    rs[ty * 4 + i][tx] = part;
  }
  __syncthreads();
  if (threadIdx.x < SIMT_BM) {
    const int b = m0 + threadIdx.x;
    if (b < a.Bp) {
      float t = 0.f;
      for (int c = 0; c < 16; ++c) t += rs[threadIdx.x][c];
      a.rs_part[((size_t)(a.k - 1) * a.nc + blockIdx.x) * a.Bp + b] = (b < a.B) ? t : 0.f;
    }
  }
}

// end of frame t: G_new = m ? ((d0-o0) delta^0 + o0 rowsum(delta^0) + ok sum_{k>=1} rowsum(delta^k)) : G_old
__global__ void k_bwd_frame_end(BwdArgs a) {
  const int b = blockIdx.x;
  const float mv = a.mvalid[(size_t)b * a.T + a.t];
  if (mv == 0.f) return;                                               // state (and its gradient) is carried
  __shared__ float sh[2];
  if (threadIdx.x == 0) {
    float r0 = 0.f, rk = 0.f;
    for (int c = 0; c < a.nc; ++c) r0 += a.rs_part[((size_t)0 * a.nc + c) * a.Bp + b];
    for (int k = 1; k < a.K; ++k)
      for (int c = 0; c < a.nc; ++c) rk += a.rs_part[((size_t)k * a.nc + c) * a.Bp + b];
    sh[0] = r0; sh[1] = rk;
  }
  __syncthreads();
  const float* d0 = a.dbuf + (size_t)((a.K - 1) & 1) * a.Bp * a.Rp;      // delta^0 lives in the slot written last
  const float add = a.o0 * sh[0] + a.ok * sh[1];
  for (int j = threadIdx.x; j < a.Rp; j += blockDim.x)
    a.G[(size_t)b * a.Rp + j] = (j < a.R) ? a.d0mo * d0[(size_t)b * a.Rp + j] + add : 0.f;
}

// persistent backward chain: the rank-1 terms of the LAST processed frame (t = 0) are still missing from G
__global__ void k_bwd_finalize_G(float* __restrict__ G, const float* __restrict__ psum2, const float* __restrict__ mvalid,
                                 int T, int Bp, int R, int Rp, int n_cta, int parity, float o0, float ok) {
  const int b = blockIdx.x;
  if (mvalid[(size_t)b * T] == 0.f) return;
  __shared__ float add;
  if (threadIdx.x == 0) {
    float s0 = 0.f, s1 = 0.f;
    const float* ps = psum2 + (size_t)parity * 256 * 2 * Bp + b;
    for (int c = 0; c < n_cta; ++c) { s0 += ps[(size_t)c * 2 * Bp]; s1 += ps[(size_t)c * 2 * Bp + Bp]; }
    add = o0 * s0 + ok * s1;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < R; j += blockDim.x) G[(size_t)b * Rp + j] += add;
}

// ---- parameter chain -------------------------------------------------------------------------------------------------
// P = sum over splits of the partial dS_k^T (Rp x Rp).  Gsym[j][i] = -(P[j][i]/alph_j + P[i][j]/alph_i) (hi, lo) and the
// per-row partial of  d log alph (S term) = sum_i P[j][i] (delta_ij - S^T[j][i]).  grid (Rp/32, Rp/32), block (32,8)
__global__ void k_gsym(const float* __restrict__ part, int splits, const float* __restrict__ ST, const float* __restrict__ alph,
                       int R, int Rp, float* __restrict__ gsym_hi, float* __restrict__ gsym_lo, float* __restrict__ rowS) {
  __shared__ float tile[32][33];
  const int j0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
  const size_t sstride = (size_t)Rp * Rp;
  // transposed block first: P[i][j] for i in i0.., j in j0..
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int i = i0 + y, j = j0 + threadIdx.x;
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += part[s * sstride + (size_t)i * Rp + j];
    tile[y][threadIdx.x] = v / alph[i];
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int j = j0 + y, i = i0 + threadIdx.x;
    float pji = 0.f;
    for (int s = 0; s < splits; ++s) pji += part[s * sstride + (size_t)j * Rp + i];
    float g = 0.f, term = 0.f;
    if (j < R && i < R) {
      g = -(pji / alph[j] + tile[threadIdx.x][y]);
      term = pji * (((i == j) ? 1.f : 0.f) - ST[(size_t)j * Rp + i]);
    }
    gsym_hi[(size_t)j * Rp + i] = g; gsym_lo[(size_t)j * Rp + i] = tf32_lo(g);
    for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
    if (threadIdx.x == 0) rowS[(size_t)j * gridDim.x + blockIdx.x] = term;
  }
}

// Final per-layer kernel, one warp per atom j:  dDt = dDt_S + dWt/alph_j ; c_j = <dDt, D^_j> ;
//   g_log_D[f][j] (+)= (dDt[j][f] - D^[j][f] c_j) D^[j][f] ;  rowacc[j][1] = -<dWt_j, Wt_j> ; rowacc[j][2] = db_j b_j
__global__ void k_param_chain(const float* __restrict__ dDt_S, int has_S, const float* __restrict__ partX, int splits,
                              int Fx, const float* __restrict__ Dt, const float* __restrict__ Wt,
                              const float* __restrict__ bias, const float* __restrict__ alph, int F, int R, int Rp, int Fp,
                              float* __restrict__ g_log_D, int accumulate, float* __restrict__ rowacc) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= R) return;
  const size_t sstride = (size_t)Rp * Fx;
  const float inv_a = 1.f / alph[j];
  float c = 0.f, tw = 0.f;
  for (int f = lane; f < F; f += 32) {
    float dw = 0.f;
    for (int s = 0; s < splits; ++s) dw += partX[s * sstride + (size_t)j * Fx + f];
    const float dd = (has_S ? dDt_S[(size_t)j * Fp + f] : 0.f) + dw * inv_a;
    c = fmaf(dd, Dt[(size_t)j * Fp + f], c);
    tw = fmaf(dw, Wt[(size_t)j * Fp + f], tw);
  }
  for (int o = 16; o > 0; o >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, o); tw += __shfl_xor_sync(0xffffffffu, tw, o); }
  for (int f = lane; f < F; f += 32) {
    float dw = 0.f;
    for (int s = 0; s < splits; ++s) dw += partX[s * sstride + (size_t)j * Fx + f];
    const float d = Dt[(size_t)j * Fp + f];
    const float dd = (has_S ? dDt_S[(size_t)j * Fp + f] : 0.f) + dw * inv_a;
    const float g = (dd - d * c) * d;
    float* dst = g_log_D + (size_t)f * R + j;
    *dst = accumulate ? (*dst + g) : g;
  }
  if (lane == 0) {
    float db = 0.f;
    for (int s = 0; s < splits; ++s) db += partX[s * sstride + (size_t)j * Fx + F];
    rowacc[(size_t)j * 4 + 1] = -tw;
    rowacc[(size_t)j * 4 + 2] = db * bias[j];
  }
}

// g_log_alph[k][...] and g_log_lam1[k] from the per-row partials (single block, fixed order)
__global__ void k_scalar_grads(const float* __restrict__ rowacc, const float* __restrict__ rowS, int nS, int R, int alph_dim,
                               float* __restrict__ g_log_alph, float* __restrict__ g_log_lam1, int acc_alph, int acc_lam) {
  __shared__ float ra[256], rl[256];
  float a = 0.f, l = 0.f;
  for (int j = threadIdx.x; j < R; j += 256) {
    float sterm = 0.f;
    for (int c = 0; c < nS; ++c) sterm += rowS[(size_t)j * nS + c];
    const float ga = sterm + rowacc[(size_t)j * 4 + 1] - rowacc[(size_t)j * 4 + 2];
    if (alph_dim > 1) g_log_alph[j] = acc_alph ? g_log_alph[j] + ga : ga;
    a += ga;
    l += rowacc[(size_t)j * 4 + 2];
  }
  ra[threadIdx.x] = a; rl[threadIdx.x] = l;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) { ra[threadIdx.x] += ra[threadIdx.x + o]; rl[threadIdx.x] += rl[threadIdx.x + o]; } __syncthreads(); }
  if (threadIdx.x == 0) {
    if (alph_dim == 1) g_log_alph[0] = acc_alph ? g_log_alph[0] + ra[0] : ra[0];
    g_log_lam1[0] = acc_lam ? g_log_lam1[0] + rl[0] : rl[0];
  }
}

// recon kernels: g_k[j][f] = exp(k[j][f]) * sum_frames H[frame][j] dS'[frame][f]   (dEc from the GEMM, Rp x Fp each)
__global__ void k_recon_grads(const float* __restrict__ dEc, const float* __restrict__ EcB, int F, int R, int r, int Rp,
                              int Fp, float* __restrict__ g_kc, float* __restrict__ g_kn) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * F) return;
  const int j = idx / F, f = idx % F;
  if (j < r) g_kc[(size_t)j * F + f] = dEc[(size_t)j * Fp + f] * EcB[(size_t)j * 2 * Fp + f];
  else g_kn[(size_t)(j - r) * F + f] = dEc[(size_t)Rp * Fp + (size_t)j * Fp + f] * EcB[(size_t)j * 2 * Fp + Fp + f];
}

// g_log_h0[j] = sigmoid(log_h0[j]) * sum_b G[b][j]
__global__ void k_h0_grad(const float* __restrict__ G, const float* __restrict__ log_h0, int B, int R, int Rp,
                          float* __restrict__ g_log_h0) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= R) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += G[(size_t)b * Rp + j];
  g_log_h0[j] = s / (1.f + expf(-log_h0[j]));
}

static int gemm(const drnmf_handle* h, GemmEpi epi, const GemmArgs& a, cudaStream_t st) {
  return h->impl == DRNMF_IMPL_SIMT ? launch_gemm_simt(epi, a, st) : launch_gemm_tc(epi, a, st);
}

// Keras 2.0.4 Adam on a flat parameter segment, fused with the 1/frames normalisation of the gradient and the
// trainable mask:  g' = g * gscale ; m = b1 m + (1-b1) g' ; v = b2 v + (1-b2) g'^2 ; p -= lr_t m / (sqrt(v) + eps).
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       const uint8_t* __restrict__ trainable, size_t n, float lr_t, float b1, float b2, float eps, float gscale) {
  const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i0 + 3 < n) {
    const float4 g4 = *reinterpret_cast<const float4*>(g + i0);
    float4 p4 = *reinterpret_cast<const float4*>(p + i0), m4 = *reinterpret_cast<const float4*>(m + i0), v4 = *reinterpret_cast<const float4*>(v + i0);
    const uchar4 t4 = trainable ? *reinterpret_cast<const uchar4*>(trainable + i0) : make_uchar4(1, 1, 1, 1);
    auto upd = [&](float& pp, float gg, float& mm, float& vv, unsigned char tt) {
      if (!tt) return;
      gg *= gscale;
      mm = b1 * mm + (1.f - b1) * gg; vv = b2 * vv + (1.f - b2) * gg * gg;
      pp -= lr_t * mm / (sqrtf(vv) + eps);
    };
    upd(p4.x, g4.x, m4.x, v4.x, t4.x); upd(p4.y, g4.y, m4.y, v4.y, t4.y); upd(p4.z, g4.z, m4.z, v4.z, t4.z); upd(p4.w, g4.w, m4.w, v4.w, t4.w);
    *reinterpret_cast<float4*>(p + i0) = p4; *reinterpret_cast<float4*>(m + i0) = m4; *reinterpret_cast<float4*>(v + i0) = v4;
  } else {
    for (size_t i = i0; i < n; ++i) {
      if (trainable && !trainable[i]) continue;
      const float gg = g[i] * gscale;
      m[i] = b1 * m[i] + (1.f - b1) * gg; v[i] = b2 * v[i] + (1.f - b2) * gg * gg;
      p[i] -= lr_t * m[i] / (sqrtf(v[i]) + eps);
    }
  }
}

int launch_adam(float* p, const float* g, float* m, float* v, const uint8_t* trainable, size_t n, float lr_t, float b1, float b2,
                float eps, float gscale, cudaStream_t st) {
  if (n == 0) return DRNMF_OK;
  DRNMF_CHECK(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (!trainable || (reinterpret_cast<uintptr_t>(trainable) & 3) == 0),
              "drnmf_adam_step: buffers must be 16-byte aligned (mask 4-byte)");
  const size_t nthreads = (n + 3) / 4;
  k_adam<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(p, g, m, v, trainable, n, lr_t, b1, b2, eps, gscale);
  count_launch();
  DRNMF_CUDA(cudaGetLastError());
  return DRNMF_OK;
}

int train_loss_and_grads(drnmf_handle* h, const float* x, const float* y, int B, int T, float mask_value, float* g_log_D,
                         float* g_log_alph, float* g_log_lam1, float* g_log_h0, float* g_k_clean, float* g_k_noise,
                         double* loss_host, float* irm_out, void* ws, size_t ws_bytes, cudaStream_t st,
                         drnmf_layer_fn layer_cb, void* cb_user) {
  DRNMF_CHECK(h->uk_d == h->uk_o, "training assumes U_k = c*11^T for k >= 1 (what build_alt creates)");
  TrainWs w = carve_train(h, B, T, ws);
  if (ws_bytes < w.bytes) { set_error("training workspace too small: need %zu bytes, got %zu", w.bytes, ws_bytes); return DRNMF_ERR_WORKSPACE; }
  const int K = h->K, R = h->R, Rp = h->Rp, F = h->F, Fp = h->Fp, r = h->r, BT = B * T, Bp = w.fwd.Bp;
  const size_t TB = (size_t)w.TB;
  const dim3 tb(32, 8);
  int rc;
  // ---------------- forward with stored activations ----------------
  DRNMF_CUDA(cudaMemsetAsync(w.fwd.actT_hi, 0, (size_t)K * Rp * TB * 4, st));     // padded utterances / rows stay zero
  DRNMF_CUDA(cudaMemsetAsync(w.fwd.actT_lo, 0, (size_t)K * Rp * TB * 4, st));
  // masking, input projection (pipelined under the recurrence like drnmf_forward) and the recurrence with stored activations;
  // the forward workspace is the head of this call's workspace; the device error word is read at the end of the step
  if ((rc = forward_core(h, x, nullptr, B, T, mask_value, nullptr, nullptr, ws, ws_bytes, (void*)st, nullptr, nullptr,
                         w.fwd.actT_hi, w.fwd.actT_lo, false))) return rc;
  const int xp_tmajor = h->last_fwd_tmajor ? 1 : 0;          // row order of xp (the transposition below reads it)
  {
    GemmArgs a{};
    a.A_hi = w.fwd.Hp_hi; a.A_lo = w.fwd.Hp_lo; a.lda = Rp;
    a.B_hi = h->EcT_hi; a.B_lo = h->EcT_lo; a.B2_hi = h->EnT_hi; a.B2_lo = h->EnT_lo; a.ldb = Rp;
    a.M = BT; a.N = F; a.Kd = Rp; a.C = irm_out ? irm_out : w.irm; a.ldc = F; a.M_valid = BT; a.N_valid = F;
    a.C_S = w.S; a.C_N = w.N; a.square = (h->flags & DRNMF_FLAG_SQUARE_IRM) ? 1 : 0;
    if ((rc = gemm(h, EPI_RECON, a, st))) return rc;
  }
  // ---------------- loss and head gradient ----------------
  const int sq = (h->flags & DRNMF_FLAG_SQUARE_IRM) ? 1 : 0;
  DRNMF_CUDA(cudaMemsetAsync(w.dSNT_hi, 0, (size_t)2 * Fp * TB * 4, st));
  DRNMF_CUDA(cudaMemsetAsync(w.dSNT_lo, 0, (size_t)2 * Fp * TB * 4, st));
  const int nblk = (BT + 7) / 8;
  k_loss_head<<<nblk, 256, 0, st>>>(x, y, w.S, w.N, w.fwd.mvalid, B, T, Bp, F, Fp, sq, h->loss_kind, w.dSN_hi, w.dSN_lo, w.dSNT_hi, w.dSNT_lo, w.loss_part);
  k_reduce_pairs<<<1, 256, 0, st>>>(w.loss_part, nblk, w.scal);
  count_launch(2);
  {   // dH = [dS | dN] . EcB^T   (BT x Rp)
    GemmArgs a{};
    a.A_hi = w.dSN_hi; a.A_lo = w.dSN_lo; a.lda = 2 * Fp;
    a.B_hi = h->EcB_hi; a.B_lo = h->EcB_lo; a.ldb = 2 * Fp;
    a.M = BT; a.N = Rp; a.Kd = 2 * Fp; a.C = w.dH; a.ldc = Rp; a.M_valid = BT; a.N_valid = Rp;
    if ((rc = gemm(h, EPI_STORE, a, st))) return rc;
  }
  if (h->loss_kind == 1) {   // + lam1 * (R/F) * mean_j |H_j| per frame and its gradient into dH
    k_l1_head<<<nblk, 256, 0, st>>>(w.fwd.Hp_hi, w.fwd.mvalid, BT, R, Rp, h->loss_lam1 / (float)F, w.dH, w.loss_part);
    k_reduce_pairs<<<1, 256, 0, st>>>(w.loss_part, nblk, w.scal + 2);
    k_add_scal<<<1, 1, 0, st>>>(w.scal);
    count_launch(3);
  }
  k_xT<<<dim3((unsigned)(TB / 32), w.Fx / 32), tb, 0, st>>>(w.fwd.xp_hi, B, T, Bp, F, Fp, w.Fx, w.xT_hi, w.xT_lo, xp_tmajor);
  count_launch();
  // ---------------- backward chain ----------------
  DRNMF_CUDA(cudaMemsetAsync(w.deltaT_hi, 0, (size_t)K * Rp * TB * 4, st));
  DRNMF_CUDA(cudaMemsetAsync(w.deltaT_lo, 0, (size_t)K * Rp * TB * 4, st));
  DRNMF_CUDA(cudaMemsetAsync(w.G, 0, (size_t)Bp * Rp * 4, st));
  DRNMF_CUDA(cudaMemsetAsync(w.dbuf, 0, (size_t)2 * Bp * Rp * 4, st));
  BwdArgs ba;
  ba.actT = w.fwd.actT_hi; ba.mvalid = w.fwd.mvalid; ba.dH = w.dH; ba.deltaT_hi = w.deltaT_hi; ba.deltaT_lo = w.deltaT_lo;
  ba.G = w.G; ba.dbuf = w.dbuf; ba.rs_part = w.rs_part; ba.B = B; ba.Bp = Bp; ba.T = T; ba.K = K; ba.R = R; ba.Rp = Rp; ba.TB = TB;
  ba.d0mo = h->u0_d - h->u0_o; ba.o0 = h->u0_o; ba.ok = h->uk_o; ba.nc = Rp / SIMT_BN;
  ba.alph = (h->alph_dim > 1) ? h->alph : nullptr;
  const dim3 gstep(Rp / SIMT_BN, (B + SIMT_BM - 1) / SIMT_BM);
  int bwd_rc = 1;
  // Weight gradients pipelined under the backward chain.  The chain occupies 64 SMs and walks the frames from the last
  // to the first; the weight-gradient GEMMs contract over TIME-major frames, so their split-K blocks over the late
  // frames are complete long before the chain ends.  The upper half of the split-K blocks of every layer is launched
  // on a second stream once the chain reports half of the frames done, the next quarter at three quarters, the rest
  // after the chain; the per-layer kernels then sum the partials in block order as before (same numbers as the
  // serial order).
  // DRNMF_TRAIN_OVERLAP=0 = serial; off when launches are serialised or stream memory operations are missing.
  bool want_pipe = false, piped = false;
  {
    const char* e = getenv("DRNMF_TRAIN_OVERLAP");
    const char* lb = getenv("CUDA_LAUNCH_BLOCKING");
    want_pipe = !(e && !strcmp(e, "0")) && h->impl != DRNMF_IMPL_SIMT && K >= 2 && w.partS_all && T >= 32 &&
                !(lb && atoi(lb) != 0) && !getenv("CUDA_INJECTION64_PATH") && !getenv("NV_NSIGHT_INJECTION_PORT_BASE") &&
                stream_wait_geq(nullptr, nullptr, 0) == 0;
    if (want_pipe && !h->hi_ready) {
      int least = 0, greatest = 0;
      DRNMF_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      DRNMF_CUDA(cudaStreamCreateWithPriority(&h->hi, cudaStreamNonBlocking, greatest));
      DRNMF_CUDA(cudaEventCreateWithFlags(&h->ev_ov[0], cudaEventDisableTiming));
      DRNMF_CUDA(cudaEventCreateWithFlags(&h->ev_ov[1], cudaEventDisableTiming));
      h->hi_ready = true;
    }
  }
  h->last_bwd_impl = 1;
  {
    bool rec_simt = (h->impl == DRNMF_IMPL_SIMT);
    const char* e = getenv("DRNMF_RECURRENT");
    if (e && !strcmp(e, "simt")) rec_simt = true;
    if (!rec_simt) {
      // K == 1 has no layer-to-layer product at all: the per-frame kernels below are the whole chain
      if (want_pipe) {
        DRNMF_CUDA(cudaMemsetAsync(w.progress, 0, 4, st));
        DRNMF_CUDA(cudaEventRecord(h->ev_ov[0], st));           // everything the early GEMMs read besides the deltas is complete
      }
      bool prog_ok = false;
      bwd_rc = (K < 2) ? 1 : launch_recurrent_bwd_tc(h, w.fwd, B, T, w.dH, w.deltaT_hi, w.deltaT_lo, w.G, w.psum2, st,
                                                     want_pipe ? w.progress : nullptr, &prog_ok);
      piped = want_pipe && bwd_rc == 0 && prog_ok;
      if (bwd_rc == 1 && K >= 2) return DRNMF_ERR_INVALID;      // no silent CUDA-core fallback (error text set by the planner)
      if (bwd_rc != 0 && bwd_rc != 1) return bwd_rc;
      if (bwd_rc == 0) {
        h->last_bwd_impl = 0;
        const int KSx = h->bwd_cfg[1], MTx = h->bwd_cfg[2];
        k_bwd_finalize_G<<<B, 128, 0, st>>>(w.G, w.psum2, w.fwd.mvalid, T, Bp, R, Rp, KSx * MTx, (T - 1) & 1, h->u0_o, h->uk_o);
        count_launch();
      }
    }
  }
  for (int t = T - 1; t >= 0 && bwd_rc == 1; --t) {
    ba.t = t; ba.k = K - 1; ba.ST = nullptr;
    k_bwd_frame_begin<<<B, 256, 0, st>>>(ba);
    for (int k = K - 1; k >= 1; --k) {
      ba.k = k; ba.ST = h->ST_hi + (size_t)(k - 1) * Rp * Rp;
      k_bwd_step<<<gstep, SIMT_THREADS, 0, st>>>(ba);
    }
    k_bwd_frame_end<<<B, 256, 0, st>>>(ba);
    count_launch(K + 1);
  }
  DRNMF_CUDA(cudaGetLastError());
  k_h0_grad<<<(R + 127) / 128, 128, 0, st>>>(w.G, h->log_h0, B, R, Rp, g_log_h0);
  count_launch();
  // ---------------- recurrence-free weight gradients + parameter chain, layer by layer ----------------
  const bool tied_D = (h->n_log_D == 1), tied_a = (h->n_log_alph == 1), tied_l = (h->n_log_lam1 == 1);
  // split-K blocks [z0, z0 + nz) of layer k's two weight-gradient GEMMs (nz = 0: all blocks, into the shared buffer)
  auto gemm_dS = [&](int k, float* out, int z0, int nz, cudaStream_t s_) {
    GemmArgs a{};   // partial dS_k^T[j][i] = sum_frames delta^k[j][frame] act^{k-1}[i][frame]
    a.A_hi = w.deltaT_hi + (size_t)k * Rp * TB; a.A_lo = w.deltaT_lo + (size_t)k * Rp * TB; a.lda = (int)TB;
    a.B_hi = w.fwd.actT_hi + (size_t)(k - 1) * Rp * TB; a.B_lo = w.fwd.actT_lo + (size_t)(k - 1) * Rp * TB; a.ldb = (int)TB;
    a.M = Rp; a.N = Rp; a.Kd = (int)TB; a.C = out; a.ldc = Rp; a.M_valid = Rp; a.N_valid = Rp;
    a.splits = w.splits_w; a.split_stride = (size_t)Rp * Rp; a.split_z0 = z0; a.split_nz = nz;
    return gemm(h, EPI_STORE, a, s_);
  };
  auto gemm_dX = [&](int k, float* out, int z0, int nz, cudaStream_t s_) {
    GemmArgs a{};   // partial [dWt_k | db_k][j][f] = sum_frames delta^k[j][frame] [x~ ; 1][f][frame]
    a.A_hi = w.deltaT_hi + (size_t)k * Rp * TB; a.A_lo = w.deltaT_lo + (size_t)k * Rp * TB; a.lda = (int)TB;
    a.B_hi = w.xT_hi; a.B_lo = w.xT_lo; a.ldb = (int)TB;
    a.M = Rp; a.N = w.Fx; a.Kd = (int)TB; a.C = out; a.ldc = w.Fx; a.M_valid = Rp; a.N_valid = w.Fx;
    a.splits = w.splits_x; a.split_stride = (size_t)Rp * w.Fx; a.split_z0 = z0; a.split_nz = nz;
    return gemm(h, EPI_STORE, a, s_);
  };
  const size_t strideS = (size_t)w.splits_w * Rp * Rp, strideX = (size_t)w.splits_x * Rp * (size_t)w.Fx;
  int z_late = w.splits_w;                                   // blocks [0, z_late) remain for the serial phase
  if (piped) {
    // block z covers the K columns [z * cols, (z + 1) * cols) of the time-major frames, i.e. the frames >= z * cols / Bp;
    // they are complete once the chain has processed T - floor(z * cols / Bp) frames
    const int kb_chunk = ((int)(TB / 32) + w.splits_w - 1) / w.splits_w;
    const long long cols = (long long)kb_chunk * 32;
    DRNMF_CUDA(cudaStreamWaitEvent(h->hi, h->ev_ov[0], 0));
    const int S = w.splits_w;                                  // 8 blocks: {4..7}, {2,3}, then {0,1}; 4 blocks: {2,3}, {1}, then {0}
    const int phase_z0[2] = {S / 2, S / 4}, phase_nz[2] = {S - S / 2, S / 2 - S / 4};
    for (int ph = 0; ph < 2; ++ph) {
      const int t_min = (int)((phase_z0[ph] * cols) / Bp);
      const unsigned int need = (unsigned int)(T - t_min);
      if (need >= (unsigned int)T) break;                      // (never for 8 blocks; guards the wait below)
      if (stream_wait_geq(h->hi, w.progress, need)) { set_error("cuStreamWaitValue32 failed"); return DRNMF_ERR_CUDA; }
      for (int k = 0; k < K; ++k) {
        if (k >= 1 && (rc = gemm_dS(k, w.partS_all + (size_t)k * strideS, phase_z0[ph], phase_nz[ph], h->hi))) return rc;
        if ((rc = gemm_dX(k, w.partX_all + (size_t)k * strideX, phase_z0[ph], phase_nz[ph], h->hi))) return rc;
      }
      z_late = phase_z0[ph];
    }
    DRNMF_CUDA(cudaEventRecord(h->ev_ov[1], h->hi));
    DRNMF_CUDA(cudaStreamWaitEvent(st, h->ev_ov[1], 0));        // (st is behind the chain here: the wait costs nothing)
  }
  for (int k = 0; k < K; ++k) {
    DRNMF_CUDA(cudaMemsetAsync(w.rowacc, 0, (size_t)Rp * 4 * 4, st));
    float* partS = piped ? w.partS_all + (size_t)k * strideS : w.part;
    if (k >= 1) {
      if ((rc = piped ? gemm_dS(k, partS, 0, z_late, st) : gemm_dS(k, partS, 0, 0, st))) return rc;
      k_gsym<<<dim3(Rp / 32, Rp / 32), tb, 0, st>>>(partS, w.splits_w, h->ST_hi + (size_t)(k - 1) * Rp * Rp, h->alph + (size_t)k * Rp,
                                                     R, Rp, w.gsym_hi, w.gsym_lo, w.rowS);
      count_launch();
      GemmArgs g{};   // dDt_S[j][f] = sum_i Gsym[j][i] D^[f][i]
      g.A_hi = w.gsym_hi; g.A_lo = w.gsym_lo; g.lda = Rp;
      g.B_hi = h->Dm_hi + (size_t)k * Fp * Rp; g.B_lo = h->Dm_lo + (size_t)k * Fp * Rp; g.ldb = Rp;
      g.M = Rp; g.N = Fp; g.Kd = Rp; g.C = w.dDt; g.ldc = Fp; g.M_valid = Rp; g.N_valid = Fp;
      if ((rc = gemm(h, EPI_STORE, g, st))) return rc;
    }
    float* partX = piped ? w.partX_all + (size_t)k * strideX : w.part;
    if ((rc = piped ? gemm_dX(k, partX, 0, z_late, st) : gemm_dX(k, partX, 0, 0, st))) return rc;
    float* gD = g_log_D + (tied_D ? 0 : (size_t)k * F * R);
    k_param_chain<<<(R + 7) / 8, 256, 0, st>>>(w.dDt, k >= 1, partX, w.splits_x, w.Fx, h->Dt_hi + (size_t)k * Rp * Fp,
                                                h->Wt_hi + (size_t)k * Rp * Fp, h->bias + (size_t)k * Rp, h->alph + (size_t)k * Rp,
                                                F, R, Rp, Fp, gD, (tied_D && k > 0) ? 1 : 0, w.rowacc);
    k_scalar_grads<<<1, 256, 0, st>>>(w.rowacc, w.rowS, k >= 1 ? Rp / 32 : 0, R, h->alph_dim, g_log_alph + (tied_a ? 0 : (size_t)k * h->alph_dim), g_log_lam1 + (tied_l ? 0 : k),
                                      (tied_a && k > 0) ? 1 : 0, (tied_l && k > 0) ? 1 : 0);
    count_launch(2);
    // the gradients of layer k are enqueued: a data-parallel caller starts their all-reduce now, under the GEMMs of
    // the layers that follow (enhance.py:1152 trains through Keras; bucketing is this build's)
    if (layer_cb && (rc = layer_cb(cb_user, k, (void*)st))) { set_error("drnmf_loss_and_grads: layer callback failed for layer %d (%d)", k, rc); return DRNMF_ERR_INVALID; }
  }
  {   // recon kernels: dEc = act^{K-1}(time-major)^T-rows . dS'^T-rows  and the same with dN'
    for (int which = 0; which < 2; ++which) {
      GemmArgs a{};
      a.A_hi = w.fwd.actT_hi + (size_t)(K - 1) * Rp * TB; a.A_lo = w.fwd.actT_lo + (size_t)(K - 1) * Rp * TB; a.lda = (int)TB;
      a.B_hi = w.dSNT_hi + (size_t)which * Fp * TB; a.B_lo = w.dSNT_lo + (size_t)which * Fp * TB; a.ldb = (int)TB;
      a.M = Rp; a.N = Fp; a.Kd = (int)TB; a.C = w.dEc + (size_t)which * Rp * Fp; a.ldc = Fp; a.M_valid = Rp; a.N_valid = Fp;
      if ((rc = gemm(h, EPI_STORE, a, st))) return rc;
    }
    k_recon_grads<<<(R * F + 255) / 256, 256, 0, st>>>(w.dEc, h->EcB_hi, F, R, r, Rp, Fp, g_k_clean, g_k_noise);
    count_launch();
  }
  DRNMF_CUDA(cudaMemcpyAsync(loss_host, w.scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  DRNMF_CUDA(cudaStreamSynchronize(st));
  DRNMF_CUDA(cudaGetLastError());
  return DRNMF_OK;
}

// ---- flag_return_all_hidden (custom_layers.py:178-181, 371-374): the concatenated hidden vectors of all K layers ----------
// Keras' masked scan carries the whole output vector over masked frames (zeros before the first valid one).
// actT: K x Rp x (T*Bp) time-major activations of the forward pass.  One thread per (b, k, j), sequential over t.
__global__ void k_gather_all_hidden(const float* __restrict__ actT, const float* __restrict__ mvalid, int B, int T, int Bp, int K,
                                    int R, int Rp, float* __restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * K * R) return;
  const int j = (int)(idx % R), k = (int)((idx / R) % K), b = (int)(idx / ((size_t)R * K));
  const float* src = actT + ((size_t)k * Rp + j) * ((size_t)T * Bp) + b;
  float prev = 0.f;
  for (int t = 0; t < T; ++t) {
    if (mvalid[(size_t)b * T + t] != 0.f) prev = src[(size_t)t * Bp];
    out[((size_t)b * T + t) * ((size_t)K * R) + (size_t)k * R + j] = prev;
  }
}

int forward_all_hidden(drnmf_handle* h, const float* x, int B, int T, float mask_value, float* H_all, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
  TrainWs w = carve_train(h, B, T, ws);
  if (ws_bytes < w.bytes) { set_error("workspace too small: need %zu bytes, got %zu", w.bytes, ws_bytes); return DRNMF_ERR_WORKSPACE; }
  const int K = h->K, Rp = h->Rp;
  const size_t TB = (size_t)w.TB;
  int rc;
  DRNMF_CUDA(cudaMemsetAsync(w.fwd.actT_hi, 0, (size_t)K * Rp * TB * 4, st));
  DRNMF_CUDA(cudaMemsetAsync(w.fwd.actT_lo, 0, (size_t)K * Rp * TB * 4, st));
  if ((rc = forward_core(h, x, nullptr, B, T, mask_value, nullptr, nullptr, ws, ws_bytes, (void*)st, nullptr, nullptr,
                         w.fwd.actT_hi, w.fwd.actT_lo, false))) return rc;
  const size_t n = (size_t)B * K * h->R;
  k_gather_all_hidden<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.fwd.actT_hi, w.fwd.mvalid, B, T, w.fwd.Bp, K, h->R, Rp, H_all);
  count_launch();
  DRNMF_CUDA(cudaGetLastError());
  return DRNMF_OK;
}

}  // namespace drnmf
