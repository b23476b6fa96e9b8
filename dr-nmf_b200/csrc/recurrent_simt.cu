// SimpleDeepRNN.step (custom_layers.py:343-375) under Keras' masked scan, CUDA-core fp32 version: one launch per
// (frame, layer).  This is the semantics-lock implementation (DRNMF_IMPL_SIMT); the product path is the persistent
// tcgen05 kernel in recurrent_tc.cu, which shares every buffer layout with this file.
#include "internal.h"
#include "gemm_simt.cuh"

namespace drnmf {

struct StepArgs {
  const float* XW;        // BT x (K*Rp)  input projections INCLUDING the bias b_k
  const float* bias;      // K x Rp
  const float* mvalid;    // BT
  float* state;           // Bp x Rp
  float* leak;            // Bp : sum_j state[b][j] of the frame being processed
  const float* h_in;      // Bp x Rp  output of layer k-1
  float* h_out_hi;        // Bp x Rp  output of layer k (hi / lo)
  float* h_out_lo;
  float *Hp_hi, *Hp_lo;   // BT x Rp
  float* H_user;          // B x T x R or null
  float *actT_hi, *actT_lo; int Bp;   // training: K x Rp x (T*Bp) activations (time-major frames) or null
  const float* ST;        // Rp x Rp for this layer
  int B, T, t, k, K, R, Rp;
  float dmo, off;         // (diag - offdiag), offdiag of U_k
};

// final-layer bookkeeping of the masked scan: out_t = m ? g : out_{t-1} (zeros before the first step),
// state = m ? g : state
__device__ __forceinline__ void finish_frame(const StepArgs& a, int b, int j, float g) {
  const size_t bt = (size_t)b * a.T + a.t;
  const bool m = a.mvalid[bt] != 0.f;
  const size_t so = (size_t)b * a.Rp + j;
  float out;
  if (m) { out = g; a.state[so] = g; }
  else out = (a.t > 0) ? a.Hp_hi[(bt - 1) * a.Rp + j] : 0.f;
  a.Hp_hi[bt * a.Rp + j] = out;
  a.Hp_lo[bt * a.Rp + j] = tf32_lo(out);
  if (a.H_user && j < a.R) a.H_user[bt * a.R + j] = out;
}

// frame start: leak[b] = sum_j state[b][j]; layer 0 (no Gram term, custom_layers.py:361-369):
//   g0 = relu(state*(d0-o0) + o0*leak + x~W_0 + b_0).  One CTA per utterance.
__global__ void k_frame_begin(StepArgs a) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float* st = a.state + (size_t)b * a.Rp;
  float s = 0.f;
  for (int j = threadIdx.x; j < a.R; j += blockDim.x) s += st[j];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) { red[0] = v; a.leak[b] = v; }
  }
  __syncthreads();
  const float leak = red[0];
  const float* xw = a.XW + ((size_t)b * a.T + a.t) * ((size_t)a.K * a.Rp);
  for (int j = threadIdx.x; j < a.Rp; j += blockDim.x) {
    float g = 0.f;
    if (j < a.R) g = fmaxf(st[j] * a.dmo + a.off * leak + xw[j], 0.f);
    if (a.actT_hi) {
      const size_t o3 = (size_t)j * ((size_t)a.T * a.Bp) + (size_t)a.t * a.Bp + b;
      a.actT_hi[o3] = g; a.actT_lo[o3] = tf32_lo(g);
    }
    if (a.K == 1) {
      finish_frame(a, b, j, g);
    } else {
      a.h_out_hi[(size_t)b * a.Rp + j] = g;
      a.h_out_lo[(size_t)b * a.Rp + j] = tf32_lo(g);
    }
  }
}

// layer k >= 1:  g^k = relu(prev.U_k + g^{k-1}.S_k + x~W_k + b_k)
__global__ void __launch_bounds__(SIMT_THREADS) k_step_simt(StepArgs a) {
  float acc[4][4], acc2[4][4];
  const int m0 = blockIdx.y * SIMT_BM, n0 = blockIdx.x * SIMT_BN;     // m = utterance, n = output atom j
  simt_tile_mainloop<false>(a.h_in, a.Rp, a.B, a.ST, nullptr, a.Rp, a.Rp, a.Rp, m0, n0, acc, acc2);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = m0 + ty * 4 + i;
    if (b >= a.B) continue;
    const float leak = a.leak[b];
    const float* xw = a.XW + ((size_t)b * a.T + a.t) * ((size_t)a.K * a.Rp) + (size_t)a.k * a.Rp;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = n0 + tx * 4 + jj;
      if (j >= a.Rp) continue;
      float g = 0.f;
      if (j < a.R) {
        float pre = a.off * leak + acc[i][jj] + xw[j];
        if (a.dmo != 0.f) pre += a.dmo * a.state[(size_t)b * a.Rp + j];
        g = fmaxf(pre, 0.f);
      }
      if (a.actT_hi) {
        const size_t o3 = ((size_t)a.k * a.Rp + j) * ((size_t)a.T * a.Bp) + (size_t)a.t * a.Bp + b;
        a.actT_hi[o3] = g; a.actT_lo[o3] = tf32_lo(g);
      }
      if (a.k == a.K - 1) {
        finish_frame(a, b, j, g);
      } else {
        a.h_out_hi[(size_t)b * a.Rp + j] = g;
        a.h_out_lo[(size_t)b * a.Rp + j] = tf32_lo(g);
      }
    }
  }
}

// state <- h0 tiled over the batch (custom_layers.py:336-341)
__global__ void k_init_state(const float* __restrict__ h0, int Bp, int Rp, float* __restrict__ state) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (size_t)Bp * Rp) state[idx] = h0[idx % Rp];
}

int launch_init_state(const drnmf_handle* h, FwdWorkspace& w, cudaStream_t st) {
  size_t n = (size_t)w.Bp * h->Rp;
  k_init_state<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->h0, w.Bp, h->Rp, w.state);
  count_launch();
  DRNMF_CUDA(cudaGetLastError());
  return DRNMF_OK;
}

int launch_recurrent_simt(drnmf_handle* h, FwdWorkspace& w, int B, int T, float* H_user, cudaStream_t st) {
  h->last_rec_impl = 1;
  int rc = launch_init_state(h, w, st);
  if (rc) return rc;
  const int Rp = h->Rp, K = h->K;
  const size_t slot = (size_t)w.Bp * Rp;
  StepArgs a;
  a.XW = w.XW; a.bias = h->bias; a.mvalid = w.mvalid; a.state = w.state; a.leak = w.leak;
  a.Hp_hi = w.Hp_hi; a.Hp_lo = w.Hp_lo; a.H_user = H_user;
  a.actT_hi = w.actT_hi; a.actT_lo = w.actT_lo; a.Bp = w.Bp;
  a.B = B; a.T = T; a.K = K; a.R = h->R; a.Rp = Rp;
  dim3 grid(Rp / SIMT_BN, (B + SIMT_BM - 1) / SIMT_BM);
  for (int t = 0; t < T; ++t) {
    a.t = t; a.k = 0; a.dmo = h->u0_d - h->u0_o; a.off = h->u0_o;
    a.h_in = nullptr; a.h_out_hi = w.hb_hi; a.h_out_lo = w.hb_lo; a.ST = nullptr;
    k_frame_begin<<<B, 256, 0, st>>>(a);
    for (int k = 1; k < K; ++k) {
      const int in_slot = (k - 1) & 1, out_slot = k & 1;
      a.k = k; a.dmo = h->uk_d - h->uk_o; a.off = h->uk_o;
      a.h_in = w.hb_hi + in_slot * slot;
      a.h_out_hi = w.hb_hi + out_slot * slot; a.h_out_lo = w.hb_lo + out_slot * slot;
      a.ST = h->ST_hi + (size_t)(k - 1) * Rp * Rp;
      k_step_simt<<<grid, SIMT_THREADS, 0, st>>>(a);
    }
    count_launch(K);
  }
  DRNMF_CUDA(cudaGetLastError());
  return DRNMF_OK;
}

}  // namespace drnmf
