// Internal declarations shared by the translation units of libdrnmf.so (not part of the C-ABI).
#pragma once
#include "common.cuh"

#define DRNMF_MAX_DEVICES 64

struct drnmf_handle {
  int F, R, K, r;      // bins, atoms (speech+noise), layers, atoms per source
  int Rp, Fp, Fq;      // R padded to 128, F padded to 32 (GEMM K dim), F padded to 128 (GEMM N dim rows)
  int flags, impl, device, num_sms;
  bool params_set;
  // derived from the parameters, owned by the handle (device memory)
  float *Dt_hi, *Dt_lo;    // K x Rp x Fp   rows j: D^_k[:, j]            (hi = fp32 value, lo = value - tf32_trunc)
  float *Wt_hi, *Wt_lo;    // K x Rp x Fp   rows j: D^_k[:, j] / alph_kj   = W_k^T      (enhance.py:183-195)
  float *bias;             // K x Rp        b_k = -lam_k / alph_kj                      (enhance.py:197-204)
  float *ST_hi, *ST_lo;    // (K-1) x Rp x Rp   S_k^T[j][i] = delta_ij - (1/alph_kj) d^_j . d^_i   (enhance.py:172-181)
  float *EcT_hi, *EcT_lo;  // Fq x Rp       exp(k_clean)^T in columns [0,r), zero elsewhere   (custom_layers.py:24)
  float *EnT_hi, *EnT_lo;  // Fq x Rp       exp(k_noise)^T in columns [r,R), zero elsewhere
  float *h0;               // Rp            softplus(log_h0)                            (custom_layers.py:206)
  // extra copies used by the backward pass
  float *Dm_hi, *Dm_lo;    // K x Fp x Rp   D^_k (rows f): B operand of dD^ = Gsym . D^
  float *EcB_hi, *EcB_lo;  // Rp x 2Fp      block-diagonal [exp(k_clean) | 0 ; 0 | exp(k_noise)]: B operand of dH = [dS|dN] . EcB^T
  float *alph;             // K x Rp        exp(log_alph) per layer and atom
  float *log_h0;           // Rp
  int n_log_D, n_log_alph, alph_dim, n_log_lam1;
  float *inv_norm;         // K x Rp scratch
  float u0_d, u0_o, uk_d, uk_o;
  int* dev_error;          // device-side error word (watchdogs / protocol violations)
  cudaEvent_t ev[5];       // stage boundaries of the last drnmf_forward (mask | projection | recurrence | recon)
  bool ev_ready, ev_valid;
  cudaStream_t hi;         // high-priority stream of the persistent recurrence when drnmf_forward pipelines the input
                           // projection under it (the rest of the projection runs on the caller's stream meanwhile)
  cudaEvent_t ev_ov[2];    // [0] first projection chunk done (caller stream), [1] recurrence done (hi stream)
  bool hi_ready;
  int plan_B, plan_ctas;   // cache: CTAs of the forward plan for batch plan_B (0 = empty)
  bool last_fwd_tmajor;    // the last forward_core laid xp / XW out time-major (pipelined order)
  bool no_overlap;         // a pipelined drnmf_forward timed out on the projection flag: keep the serial order
  cudaStream_t side;       // copy stream of drnmf_enhance_host (the complex STFT is only needed after the recurrence)
  cudaEvent_t ev_side[2];  // [0] main stream reached the call, [1] side copy done
  bool side_ready;
  int last_rec_impl;       // 0 = persistent tcgen05 kernel, 1 = SIMT per-step kernels (last drnmf_forward)
  int last_bwd_impl;       // the same for the backward chain of the last drnmf_loss_and_grads
  int rec_cfg[8];          // NB, KS, MT, ATOMS, n_tiles, WST, HST, RST of the last persistent launch
  int rec_groups;          // batch groups (grid.z) of the last persistent launch
  int bwd_cfg[8], bwd_groups;   // the same for the backward chain of the last drnmf_loss_and_grads
  int loss_kind;           // 0 = 'mse_of_masked' (enhance.py:1040-1047), 1 = SNMF pretraining cost (enhance.py:1024-1036)
  float loss_lam1;         // lam1 of the pretraining cost
};

namespace drnmf {

void count_launch(int n = 1);

// ---- workspace carving for forward ------------------------------------------------------------
struct FwdWorkspace {
  float *xp_hi, *xp_lo;     // BT x Fp     masked, zero-padded input and its tf32 remainder
  float* mvalid;            // BT          1.0 where the frame is valid (Keras Masking)
  float* XW;                // BT x (K*Rp) input projections x~_t . W_k + b_k for every layer
  float *Hp_hi, *Hp_lo;     // BT x Rp     padded output sequence (A operand of the recon GEMM)
  float *hb_hi, *hb_lo;     // 2 x Bp x Rp ping-pong hidden state between layers
  float* state;             // Bp x Rp     recurrent state (last layer of the previous valid frame)
  float* psum;              // 2 x 256 x Bp partial row sums of the state (rank-1 leak)
  float* leak;              // Bp          SIMT path: sum_j state[b][j]
  unsigned int* flags;      // device flags for the persistent kernel
  float *actT_hi, *actT_lo; // training only: K x Rp x (T*Bp) post-relu activations, time-major frames (t*Bp + b)
  // pipelined forward (drnmf_forward): xp / XW rows are TIME-major (t*B + b) so that the projection of the first xw_t0
  // frames is a contiguous row block; the rest is computed while the recurrence already runs and announced through
  // *xw_ready (the owners of the persistent kernel acquire it before they touch frame xw_t0).  0 / null = b-major, no wait.
  int xw_tmajor, xw_t0;
  unsigned int* xw_ready;   // [0] projections of the second chunk are complete, [1] resident CTAs of the persistent kernel
  size_t bytes;
  int Bp;
};
FwdWorkspace carve_forward_ws(const drnmf_handle* h, int B, int T, void* base);

// ---- prep.cu -----------------------------------------------------------------------------------
int launch_prep_params(drnmf_handle* h, const float* log_D, int n_log_D, const float* log_alph, int n_log_alph,
                       int alph_dim, const float* log_lam1, int n_log_lam1, const float* log_h0, const float* k_clean,
                       const float* k_noise, cudaStream_t st);
// B_tmajor > 0: write xp rows time-major (row t*B + b for input row b*T + t, T = BT / B); mvalid stays b-major
// (t_count > 0: only the frames [t_begin, t_begin + t_count) of every utterance)
int launch_mask_pad(const drnmf_handle* h, const float* x, int BT, float mask_value, FwdWorkspace& w, cudaStream_t st,
                    int B_tmajor = 0, int t_begin = 0, int t_count = 0);

// ---- gemm: C = A (M x K, K-major) . B^T (N x K, K-major) with a fused epilogue --------------------
enum GemmEpi { EPI_STORE = 0, EPI_GRAM = 1, EPI_RECON = 2, EPI_LAMBDA = 3, EPI_LAMBDA_B = 4, EPI_MU_H = 5 };
struct GemmArgs {
  const float *A_hi, *A_lo; int lda;     // M x Kd
  const float *B_hi, *B_lo; int ldb;     // N x Kd
  const float *B2_hi, *B2_lo;            // second B (EPI_RECON: noise dictionary)
  int M, N, Kd;                          // logical extents (tiles beyond are zero-filled)
  float *C, *C_lo; int ldc;              // outputs
  int R_valid, N_valid, M_valid;         // masks for the epilogues
  int square;                            // EPI_RECON: transform_before_irm == 'square'
  const float* bias;                     // EPI_STORE: optional per-column bias added in the epilogue (length N)
  float *C_S, *C_N;                      // EPI_RECON: optional raw reconstructions S^, N^ (same ld as C) for the backward pass
  int splits;                            // split-K: grid.z partial products, split z written at C + z*split_stride (0/1 = off)
  size_t split_stride;
  // EPI_LAMBDA (sparse NMF): C/C_lo = max(acc, flr) (M x ldc), CT/CT_lo = its transpose (N x ldct), and the squared
  // error against Vref (M x ldv) summed per CTA into div_partials[blockIdx.y * gridDim.x + blockIdx.x]
  float *CT, *CT_lo; int ldct;
  const float* Vref; int ldv;
  double* div_partials;
  float flr;
  // EPI_LAMBDA_B (beta-divergence, beta != 2; sparse_nmf_gpu.m:212-276): with L = max(acc, flr) the epilogue writes
  // P = L^(beta-1) to C/C_lo/CT/CT_lo (the operand that takes Lambda's place in the updates) and Q = Vref . L^(beta-2)
  // to Q/Q_lo/QT/QT_lo (the operand that takes V's place), and sums the beta-divergence into div_partials.
  float beta;
  float *Q, *Q_lo, *QT, *QT_lo;
  // tile-shape decision: when this launch is a row block of a larger product (pipelined projection), M_plan = that
  // product's row count, so that every block accumulates in the same order as the single launch would (bitwise equal)
  int M_plan;
  // EPI_MU_H (dual-B; sparse_nmf_gpu.m:217-228): acc = (W^T P)[r][frame], acc2 = (W^T Q)[r][frame]; the epilogue applies
  // h <- h .* acc2 ./ max(acc + mu, flr) to the rows with row_update != 0 (null = all) of H = C (R x ldc, in/out), writes
  // C_lo, the transposed copy CT / CT_lo and the tile's sum of H into div_partials (for cost = div + mu sum H)
  float mu;
  const uint8_t* row_update;
  // split-K sub-range: launch only the blocks [split_z0, split_z0 + split_nz) of the `splits` K-blocks (split_nz = 0: all).
  // The block boundaries are those of the full product, so a product can be computed in several launches as its
  // K range (time-major frames in the weight-gradient GEMMs) becomes available.
  int split_z0, split_nz;
};
int launch_gemm_simt(GemmEpi epi, const GemmArgs& a, cudaStream_t st);
int launch_gemm_tc(GemmEpi epi, const GemmArgs& a, cudaStream_t st);

// ---- recurrent.cu ----------------------------------------------------------------------------------
int launch_recurrent_simt(drnmf_handle* h, FwdWorkspace& w, int B, int T, float* H_user, cudaStream_t st);
int launch_recurrent_tc(drnmf_handle* h, FwdWorkspace& w, int B, int T, float* H_user, cudaStream_t st);
// api.cu: the forward pass with its pipelining (drnmf_forward, drnmf_enhance_host, and the forward of the training step)
typedef int (*upload_hook_fn)(void*);
int forward_core(drnmf_handle* h, const float* x, const float* x_host, int B, int T, float mask_value, float* H, float* irm,
                 void* ws, size_t ws_bytes, void* stream, upload_hook_fn after_upload = nullptr, void* hook_arg = nullptr,
                 float* actT_hi = nullptr, float* actT_lo = nullptr, bool final_check = true);
int recurrent_plan_ctas(const drnmf_handle* h, int B);   // CTAs of the forward plan for batch B (a huge value when there is none or it cannot be pipelined)
// progress (optional, device word, zeroed by the caller): the chain stores (release) the number of completely processed
// frames (t = T-1, T-2, ...) as it goes; only honoured for single-tile, single-group plans (*progress_ok says so)
int launch_recurrent_bwd_tc(drnmf_handle* h, FwdWorkspace& w, int B, int T, const float* dH, float* deltaT_hi,
                            float* deltaT_lo, float* G, float* psum2, cudaStream_t st, unsigned int* progress = nullptr,
                            bool* progress_ok = nullptr);

// ---- stft.cu -------------------------------------------------------------------------------------
// fidx: (n_utt, 2) int64 (start, end) frame indices per utterance, the reference's fidx (util.py:335-337)
int launch_stft_mag(const float* audio, const int64_t* offs, const int32_t* lens, const int64_t* fidx, int n_utt,
                    int max_frames, int N, int hop, int64_t total_frames, float* stack, float* mag, cudaStream_t st);
int launch_mask_istft(const float* stack, const float* mask, const int64_t* fidx, const int64_t* out_offs, int n_utt,
                      int max_frames, int N, int hop, int64_t total_frames, float* frames_tmp, float* out_audio,
                      cudaStream_t st);
// ---- snmf.cu ---------------------------------------------------------------------------------------
size_t snmf_workspace_bytes(int F, int n, int R, float beta = 2.f);
int snmf_mu_ed(int F, int n, int R, float beta, const float* V, float* W, float* H, const uint8_t* w_update, const uint8_t* h_update,
               int any_w_update, int any_h_update, float sparsity, int max_iter, float conv_eps, double* cost_host,
               double* div_host, int* iters_host, void* ws, size_t ws_bytes, bool simt, cudaStream_t st,
               drnmf_allreduce_fn allreduce, void* user);
size_t snmf_irm_workspace_bytes(int F, int n, int R);
int snmf_irm(int F, int n, int R, int r, const float* W, const float* H, float* irm, void* ws, size_t ws_bytes, bool simt,
             cudaStream_t st);
size_t ista_workspace_bytes(int F, int n, int R);
int ista_ed(int F, int n, int R, const float* x, const float* W, float* H, float lam1, float alph, int iters, void* ws,
            size_t ws_bytes, bool simt, cudaStream_t st);
// ---- train.cu --------------------------------------------------------------------------------------
size_t train_workspace_bytes(const drnmf_handle* h, int B, int T);
int train_loss_and_grads(drnmf_handle* h, const float* x, const float* y, int B, int T, float mask_value, float* g_log_D,
                         float* g_log_alph, float* g_log_lam1, float* g_log_h0, float* g_k_clean, float* g_k_noise,
                         double* loss_host, float* irm_out, void* ws, size_t ws_bytes, cudaStream_t st,
                         drnmf_layer_fn layer_cb, void* cb_user);
int forward_all_hidden(drnmf_handle* h, const float* x, int B, int T, float mask_value, float* H_all, void* ws, size_t ws_bytes,
                       cudaStream_t st);
int launch_adam(float* p, const float* g, float* m, float* v, const uint8_t* trainable, size_t n, float lr_t, float b1, float b2,
                float eps, float gscale, cudaStream_t st);
int launch_init_state(const drnmf_handle* h, FwdWorkspace& w, cudaStream_t st);
int gemm_device_error(cudaStream_t st);
const char* last_error();
unsigned long long launch_count();

}  // namespace drnmf
