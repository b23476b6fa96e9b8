/* libdrnmf.so -- C-ABI of the B200-native DR-NMF hot path.
 *
 * The reference (stwisdom/dr-nmf) has no FFI: its hot path sits behind Python-level interfaces (Keras layers,
 * a MATLAB subprocess, numpy/librosa helpers).  Each entry point below states the reference interface it
 * replaces (file:line relative to the reference repo); INTEGRATION.md shows the ctypes binding a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - return 0 = ok, nonzero = error (drnmf_status); drnmf_last_error() gives the message of the last failure
 *     on the calling thread.  Nothing is silently dropped (the reference constructs-and-drops its errors,
 *     snmf.py:105-106).
 *   - all `const float*` / `float*` arguments are DEVICE pointers unless the name ends in `_host`.
 *   - the caller owns every input/output/workspace buffer; the library only owns what a handle derives from the
 *     parameters (Gram matrices, normalised dictionaries).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  A handle is bound to the device that
 *     was current at creation and is not thread-safe (one handle per GPU process).
 *   - there is no CPU fallback: without an sm_100 device every compute call fails with DRNMF_ERR_NO_DEVICE.
 *   - all arithmetic is fp32 storage; contractions run on tcgen05 tensor cores as 3xTF32 (error-compensated).
 */
#ifndef DRNMF_H_
#define DRNMF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DRNMF_API __attribute__((visibility("default")))
#else
#define DRNMF_API
#endif

typedef struct drnmf_handle drnmf_handle;

enum drnmf_status {
  DRNMF_OK = 0,
  DRNMF_ERR_INVALID = 1,
  DRNMF_ERR_CUDA = 2,
  DRNMF_ERR_WORKSPACE = 3,
  DRNMF_ERR_DEVICE = 4,
  DRNMF_ERR_NO_DEVICE = 5
};

/* drnmf_create flags */
#define DRNMF_IMPL_TCGEN05 0      /* default: tcgen05/TMEM/TMA kernels                                        */
#define DRNMF_IMPL_SIMT 1         /* plain CUDA-core fp32 kernels (semantics lock / debugging; same layouts)  */
#define DRNMF_FLAG_SQUARE_IRM 16  /* transform_before_irm == 'square' (enhance.py:294-300)                    */

DRNMF_API int drnmf_version(void);
DRNMF_API const char* drnmf_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches claim) */
DRNMF_API unsigned long long drnmf_launch_count(void);

/* ---- model handle: SimpleDeepRNN + recon + mask  (custom_layers.py:104-412, enhance.py:139-317) ---------- */
DRNMF_API int drnmf_create(drnmf_handle** out, int F, int R, int K_layers, int flags);
DRNMF_API int drnmf_destroy(drnmf_handle* h);

/* Replaces SimpleDeepRNN.build + the build_alt maps (custom_layers.py:187-294, enhance.py:163-204): derives
 * D^_k, W_k = D^_k/alph_k, S_k = (I - (D^_k/alph_k)^T D^_k)^T, b_k = -lam_k/alph_k, h0 = softplus(log_h0) and
 * the recon kernels exp(k_clean), exp(k_noise) (custom_layers.py:24) on the device.
 *   log_D     (n_log_D, F, R)            n_log_D    in {1 (tied), K}
 *   log_alph  (n_log_alph, alph_dim)     n_log_alph in {1, K}, alph_dim in {1, R ('untie_alph', enhance.py:225)}
 *   log_lam1  (n_log_lam1)               n_log_lam1 in {1, K}
 *   log_h0 (R) ; k_clean, k_noise (R/2, F)  -- Keras kernel layout of DenseNonNegW (enhance.py:283,292)
 *   u0_diag/u0_off, uk_diag/uk_off: diagonal / off-diagonal value of exp(log_U1)^T and exp(log_Uk)^T
 *   (enhance.py:163-167); only that (a*I + b*11^T) structure is supported -- it is what build_alt creates. */
DRNMF_API int drnmf_set_params(drnmf_handle* h, const float* log_D, int n_log_D, const float* log_alph, int n_log_alph,
                     int alph_dim, const float* log_lam1, int n_log_lam1, const float* log_h0,
                     const float* k_clean, const float* k_noise, float u0_diag, float u0_off, float uk_diag,
                     float uk_off, void* stream);

DRNMF_API size_t drnmf_workspace_bytes(const drnmf_handle* h, int B, int T);

/* Replaces model_irm.predict_on_batch (enhance.py:1191-1193): Masking(mask_value) -> SimpleDeepRNN -> recon ->
 * divide_A_by_AplusB.  x (B,T,F) padded with mask_value; H (B,T,R) and irm (B,T,F) outputs (either may be NULL).
 * The call returns after the work is complete on `stream` (it reads back the device error word).  With at most 64
 * utterances the projection of all but the first frames is computed NEXT to the persistent recurrence: the library then
 * also uses a high-priority stream of its own, ordered after and joined back into `stream` (same results, bit for
 * bit; DRNMF_FWD_OVERLAP=0 keeps everything on `stream`; automatically off under CUDA_LAUNCH_BLOCKING or an injected
 * profiler / sanitizer, where kernels cannot overlap). */
DRNMF_API int drnmf_forward(drnmf_handle* h, const float* x, int B, int T, float mask_value, float* H, float* irm, void* ws,
                  size_t ws_bytes, void* stream);

/* Device time (ms, CUDA events on the call's stream) of the four stages of the last drnmf_forward on this handle:
 * [0] Masking + padding, [1] input-projection GEMM (only the part in front of the recurrence when the rest is pipelined
 * under it), [2] recurrence over (T x K_layers), [3] recon + mask GEMM. */
DRNMF_API int drnmf_stage_times(drnmf_handle* h, float* ms4);

/* How the recurrence of the last drnmf_forward ran: cfg9[0] = 0 persistent tcgen05 kernel / 1 SIMT per-step kernels;
 * cfg9[1..8] = batch tile NB, K-splits (cluster size) KS, M-tiles MT, 64-atom sub-chunks per K-slice, batch tiles per
 * group, weight / hidden / reduction ring depths of the persistent kernel. */
DRNMF_API int drnmf_recurrent_config(const drnmf_handle* h, int* cfg9);

/* Extended form: which = 0 the recurrence of the last drnmf_forward, 1 the backward chain of the last
 * drnmf_loss_and_grads; cfg10[0..8] as above, cfg10[9] = batch groups (independent utterance ranges that run the chain
 * on disjoint SMs; the grid is KS x MT x groups CTAs).  Replaces nothing in the reference (introspection only). */
DRNMF_API int drnmf_recurrent_config2(const drnmf_handle* h, int which, int* cfg10);

/* Test hook: writes `code` into the handle's device-side error word, as a kernel watchdog would.  The next compute
 * call reports DRNMF_ERR_DEVICE once and clears the word (errors do not latch). */
DRNMF_API int drnmf_debug_inject_error(drnmf_handle* h, int code, void* stream);

/* Debug/inspection: copy a derived tensor to a caller device buffer.  which: 0 = S_k^T (Rp x Rp, k>=1),
 * 1 = W_k^T (Rp x Fp), 2 = b_k (Rp), 3 = h0 (Rp).  Rp/Fp via drnmf_padded_dims. */
DRNMF_API int drnmf_get_derived(const drnmf_handle* h, int which, int k, float* out, void* stream);
DRNMF_API int drnmf_padded_dims(const drnmf_handle* h, int* Rp, int* Fp);

/* ---- STFT analysis / masked synthesis (util.py:171-226, :48-169; audio_dataset.py:194,22-23,267-278) -------
 * Utterance u occupies audio[offs[u] .. offs[u]+lens[u]) and frames [fidx[u][0], fidx[u][1]) of the stacks; fidx is
 * the reference's (n_utt, 2) (start, end) table (util.py:335-337) as int64.  max_frames >= max_u frames_u.
 * stft_frames(n, N, hop) = ceil(n/hop) + N/hop + 1 (util.py:184-190 + librosa center=False).
 * stack is the reference's [Re; Im] layout (2F, total_frames) row-major (util.py:351); mag is (total_frames, F)
 * row-major, i.e. already the (T,F) slices enhance.py feeds the network.  Either output may be NULL. */
DRNMF_API int drnmf_stft_frames(int nsampl, int N, int hop);
DRNMF_API int drnmf_stft_mag(const float* audio, const int64_t* offs, const int32_t* lens, const int64_t* fidx, int n_utt,
                   int max_frames, int N, int hop, int64_t total_frames, float* stack, float* mag, void* stream);
/* Replaces AudioDataset.reconstruct_x (audio_dataset.py:267-278): mask (total_frames, F) row-major (may be NULL),
 * out_audio: utterance u written at out_offs[u], length hop*(frames_u-1) - N  (istft_mc trims N both ends). */
DRNMF_API int drnmf_mask_istft(const float* stack, const float* mask, const int64_t* fidx, const int64_t* out_offs, int n_utt,
                     int max_frames, int N, int hop, int64_t total_frames, float* out_audio, void* ws,
                     size_t ws_bytes, void* stream);
DRNMF_API size_t drnmf_istft_workspace_bytes(int64_t total_frames, int N);

/* ---- end-to-end inference with HOST buffers (enhance.py:1186-1203 predict + reconstruct loop) -------------
 * x_host (B,T,F) padded magnitudes, stack_host (2F, B*T) [Re;Im] of the same utterances laid out utterance-major
 * with T frames each, frames_host[B] valid frame counts; audio_out_host (B, hop*(T-1)-N) zero beyond each
 * utterance's length.  H2D and D2H copies happen inside (pinned memory recommended). */
DRNMF_API int drnmf_enhance_host(drnmf_handle* h, const float* x_host, const float* stack_host, const int32_t* frames_host,
                       int B, int T, int N, int hop, float mask_value, float* audio_out_host, void* ws,
                       size_t ws_bytes, void* stream);
DRNMF_API size_t drnmf_enhance_workspace_bytes(const drnmf_handle* h, int B, int T, int N, int hop);

/* ---- sparse NMF multiplicative updates, Euclidean (sparseNMF/sparse_nmf_gpu.m:163-298; snmf.py:88-113) -----
 * Replaces the MATLAB subprocess of sparse_nmf_matlab_on_chunk: V (F,n), W (F,R) in/out, H (R,n) in/out, row-major
 * device arrays; w_update_host / h_update_host: HOST byte masks of length R (NULL = update all; :150-155);
 * cost_host / div_host: host arrays of max_iter doubles (objective.cost/.div, :271-281); *iters_host = iterations
 * run (convergence test of :288-296).  W's columns are normalised and H rescaled on entry (:163-166).
 * flags: DRNMF_IMPL_TCGEN05 / DRNMF_IMPL_SIMT. */
DRNMF_API int drnmf_snmf_mu_ed(int F, int n, int R, const float* V, float* W, float* H, const uint8_t* w_update_host,
                     const uint8_t* h_update_host, float sparsity, int max_iter, float conv_eps, double* cost_host,
                     double* div_host, int* iters_host, int flags, void* ws, size_t ws_bytes, void* stream);
DRNMF_API size_t drnmf_snmf_workspace_bytes(int F, int n, int R);
/* Frame-sharded (multi-GPU) variant: every rank holds a slice of the frames (columns of V and H) and a replica of W.
 * `allreduce(user, dev_buf, count, dtype, stream)` must sum `count` elements (dtype 0 = float32, 1 = float64) of
 * dev_buf in place over all ranks, ordered on `stream`, and return 0.  It is called twice per W-update (V H^T and
 * Lambda H^T, F x ceil32(R) floats each) and once per iteration for (div, mu*sum H); this is the un-chunked reference
 * algorithm (n_chunks == 1, snmf.py:82-83) with identical results on every rank.  NULL = single process. */
typedef int (*drnmf_allreduce_fn)(void* user, void* dev_buf, size_t count, int dtype, void* stream);
DRNMF_API int drnmf_snmf_mu_ed_dist(int F, int n, int R, const float* V, float* W, float* H, const uint8_t* w_update_host,
                          const uint8_t* h_update_host, float sparsity, int max_iter, float conv_eps, double* cost_host,
                          double* div_host, int* iters_host, int flags, void* ws, size_t ws_bytes, void* stream,
                          drnmf_allreduce_fn allreduce, void* user);

/* ---- sparse NMF multiplicative updates for any beta-divergence (sparseNMF/sparse_nmf_gpu.m:100-115 cf/beta selection,
 * :201-205 zero entries of V raised to its smallest positive entry, :212-226 H updates, :232-260 W updates, :266-276
 * divergence): beta = 1 is 'kl' (the solver's default), 0 is 'is', 2 is 'ed' (identical to drnmf_snmf_mu_ed), anything
 * else the generic branch.  Same arguments and conventions as drnmf_snmf_mu_ed_dist; with a non-NULL `allreduce` and
 * beta != 2 the callback is additionally called once with dtype 2 = float32 MIN over ranks (1 element: min positive V). */
DRNMF_API int drnmf_snmf_mu_beta(int F, int n, int R, float beta, const float* V, float* W, float* H,
                       const uint8_t* w_update_host, const uint8_t* h_update_host, float sparsity, int max_iter,
                       float conv_eps, double* cost_host, double* div_host, int* iters_host, int flags, void* ws,
                       size_t ws_bytes, void* stream, drnmf_allreduce_fn allreduce, void* user);
DRNMF_API size_t drnmf_snmf_beta_workspace_bytes(int F, int n, int R, float beta);

/* ---- SNMF baseline ratio mask (enhance.py:847-852): irm = S^ / (1e-9 + S^ + N^) with S^ = W[:, :r] H[:r],
 * N^ = W[:, r:] H[r:].  W (F,R), H (R,n), irm (F,n) row-major device arrays.  One dual-operand tcgen05 GEMM with the ratio
 * fused into its epilogue (the reference: two np.dot and an elementwise divide on the host). */
DRNMF_API int drnmf_snmf_irm(int F, int n, int R, int r, const float* W, const float* H, float* irm, int flags, void* ws,
                   size_t ws_bytes, void* stream);
DRNMF_API size_t drnmf_snmf_irm_workspace_bytes(int F, int n, int R);

/* ---- frame-parallel ISTA with a tied dictionary (enhance.py:402-418 `ista_ed`; defined but never called there) ---
 * x (F,n), W (F,R), H (R,n) in/out, row-major device arrays: H <- max(0, -lam1/alph + H + (1/alph) W^T (x - W H)),
 * `iters` times.  n must be a multiple of 4. */
DRNMF_API int drnmf_ista_ed(int F, int n, int R, const float* x, const float* W, float* H, float lam1, float alph,
                  int iters, int flags, void* ws, size_t ws_bytes, void* stream);
DRNMF_API size_t drnmf_ista_workspace_bytes(int F, int n, int R);

/* ---- training: loss and gradients through the unfolded layers (enhance.py:1040-1073, 1152-1157) -----------------
 * The reference trains with Keras/Theano autodiff (BPTT through scan); this is the hand-written equivalent.
 * x, y (B,T,F) padded with mask_value (y is only read at valid frames).  loss_host[0] = sum_{b,t} m * mean_f (x*irm-y)^2,
 * loss_host[1] = sum m; the training loss is loss_host[0]/loss_host[1] and every gradient returned is the gradient of
 * loss_host[0] (divide by the -- possibly all-reduced -- frame count).  Gradient shapes follow drnmf_set_params:
 * g_log_D (n_log_D,F,R), g_log_alph (n_log_alph), g_log_lam1 (n_log_lam1), g_log_h0 (R), g_k_clean/g_k_noise (R/2,F);
 * tied parameters receive the sum over layers.  irm (B,T,F) optional output.  log_alph may be one scalar per layer or one value per atom
 * (alph_dim = R, 'untie_alph'); g_log_alph then has n_log_alph * alph_dim entries. */
DRNMF_API int drnmf_loss_and_grads(drnmf_handle* h, const float* x, const float* y, int B, int T, float mask_value,
                         float* g_log_D, float* g_log_alph, float* g_log_lam1, float* g_log_h0, float* g_k_clean,
                         float* g_k_noise, double* loss_host, float* irm, void* ws, size_t ws_bytes, void* stream);
DRNMF_API size_t drnmf_train_workspace_bytes(const drnmf_handle* h, int B, int T);

/* Training objective of drnmf_loss_and_grads*: kind 0 = 'mse_of_masked' (enhance.py:1040-1047, the default);
 * kind 1 = the optional SNMF pretraining cost of enhance.py:1024-1036 (two-output model, loss weights [0.5, lam1 2r/F]):
 * per valid frame 0.5 mean_f (S^ + N^ - x)^2 + lam1 (R/F) mean_j |H_j|; `y` is ignored (the reference passes x twice). */
DRNMF_API int drnmf_set_training_loss(drnmf_handle* h, int kind, float lam1);

/* SimpleDeepRNN(flag_return_all_hidden=True) (custom_layers.py:178-181, 371-374): H_all (B,T,K*R) = the hidden vectors
 * of all K layers of every frame, concatenated layer-major; masked frames carry the previous frame's vector (Keras
 * masked scan).  Workspace: drnmf_train_workspace_bytes(h, B, T). */
DRNMF_API int drnmf_forward_all_hidden(drnmf_handle* h, const float* x, int B, int T, float mask_value, float* H_all,
                             void* ws, size_t ws_bytes, void* stream);

/* Data-parallel variant: `layer_ready(user, k, stream)` is called on the host right after the kernels that produce the
 * gradients of layer k (g_log_D[k], g_log_alph[k], g_log_lam1[k]; k = 0 .. K-1 in order) have been enqueued on `stream`,
 * so that the caller can start that layer's gradient all-reduce (ordered after an event it records on `stream`) under the
 * weight-gradient GEMMs of the layers that follow.  Tied parameters accumulate over the layers: their gradient is
 * complete at k = K-1.  g_log_h0, g_k_clean, g_k_noise are complete when the call returns.  Return 0 from the callback;
 * NULL = no callback.  The reference trains through Keras' model.fit (enhance.py:1152-1157) on one device. */
typedef int (*drnmf_layer_fn)(void* user, int layer, void* stream);
DRNMF_API int drnmf_loss_and_grads_cb(drnmf_handle* h, const float* x, const float* y, int B, int T, float mask_value,
                            float* g_log_D, float* g_log_alph, float* g_log_lam1, float* g_log_h0, float* g_k_clean,
                            float* g_k_noise, double* loss_host, float* irm, void* ws, size_t ws_bytes, void* stream,
                            drnmf_layer_fn layer_ready, void* user);

/* Keras 2.0.4 Adam update (optimizers.py: lr_t = lr sqrt(1-b2^t)/(1-b1^t), passed in by the caller) fused with the
 * gradient normalisation on `n` consecutive floats:  g' = grads * grad_scale ; m = b1 m + (1-b1) g' ;
 * v = b2 v + (1-b2) g'^2 ; params -= lr_t m / (sqrt(v) + epsilon).  `trainable` (n bytes, may be NULL = all) freezes
 * entries whose byte is 0 (enhance.py:239-248 `params_trainable`).  All device pointers, 16-byte aligned. */
DRNMF_API int drnmf_adam_step(float* params, const float* grads, float* m, float* v, const uint8_t* trainable, size_t n,
                    float lr_t, float beta_1, float beta_2, float epsilon, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRNMF_H_ */
