"""numpy restatement of the DR-NMF hot path of stwisdom/dr-nmf  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Every function cites the reference file:line it follows (paths relative to the reference repo).
``dtype=np.float64`` gives the ground truth used by the parity tests, ``dtype=np.float32`` is the
"what the reference's own float32 graph would compute" mode that bench.py times as ``cpu_baseline``.

Pinning: see oracle/__init__.py and oracle/pin_reference.py.  Pieces marked [unpinned] restate
third-party behaviour (Keras 2.0.4 / librosa 0.5.1 / MATLAB / BSS-Eval) that is not vendored in the reference.
"""
from __future__ import annotations

import copy
import math

import numpy as np

__all__ = [
    "EPS", "alt_params_init", "layer_weights", "structured_U", "rnn_forward", "output_head",
    "drnmf_forward", "training_loss", "ista_ed", "sparse_nmf_ed", "sparse_nmf_beta", "sparse_nmf_chunked", "train_snmf",
    "snmf_irm", "sqrt_hann", "stft_mc", "stack_reim", "magnitude", "istft_no_div", "istft_mc",
    "reconstruct_x", "wav_quantize", "sdr_db", "masked_seqs_to_frames", "param_count_notebook",
    "reshape_and_pad_stacks", "clip_x_to_y", "get_mask_value", "data_transform", "snr_db", "snmf_savefile_stem",
]

EPS = 1e-7  # the reference's ubiquitous guard (enhance.py:147, custom_layers.py:44)


# --------------------------------------------------------------------------------------------
# A.1  parameterisation (enhance.py:139-206 build_alt, :209-235 build_unfolded_snmf)
# --------------------------------------------------------------------------------------------
def alt_params_init(W, alph, lam1, K_layers, untie_alph=False, rng=None):
    """Initial parameter set of an *untied* DR-NMF model, as build_unfolded_snmf creates it.

    enhance.py:219-226 (float32 constants, optional vector alph), :147 (log(1e-7 + .) parameterisation),
    :151-159 (every untied layer starts from the same array), :282-292 (recon kernels from the un-normalised W),
    custom_layers.py:202-206 (log_h0 ~ Keras 'uniform' = U(-0.05, 0.05) [unpinned: Keras default range]).
    Returns a dict of float32 arrays: log_D (K,F,R), log_alph (K,) or (K,R), log_lam1 (K,), log_U1 (R,R),
    log_Uk (R,R), log_h0 (R,), k_clean (r,F), k_noise (r,F).
    """
    W = np.float32(W)
    F, R = W.shape
    r = R // 2
    eps32 = np.float32(EPS)
    log_D = np.log(EPS + W).astype(np.float32)              # np.log(1e-7+params['W'])   (:147)
    a = np.float32(alph)
    if untie_alph:
        a = a * np.ones((R,), dtype=np.float32)             # :225-226
    log_alph = np.log(eps32 + a).astype(np.float32)
    log_lam1 = np.log(eps32 + np.float32(lam1)).astype(np.float32)
    log_U1 = np.log(eps32 + np.eye(R, dtype=np.float32))    # :147 with U1 = I (:220)
    log_Uk = np.log(eps32 + np.zeros((R, R), dtype=np.float32))
    if rng is None:
        rng = np.random.default_rng(0)
    log_h0 = rng.uniform(-0.05, 0.05, size=(R,)).astype(np.float32)
    return {
        "log_D": np.repeat(log_D[None], K_layers, axis=0),
        "log_alph": np.repeat(np.asarray(log_alph)[None], K_layers, axis=0),
        "log_lam1": np.repeat(np.asarray(log_lam1)[None], K_layers, axis=0),
        "log_U1": log_U1, "log_Uk": log_Uk, "log_h0": log_h0,
        "k_clean": np.log(EPS + W[:, :r]).astype(np.float32).T.copy(),   # log_W_clean.T  (:282-283)
        "k_noise": np.log(EPS + W[:, r:]).astype(np.float32).T.copy(),   # :291-292
    }


def layer_weights(p, k, dtype=np.float64):
    """(W_k, S_k, b_k) of layer k exactly as the build_alt lambdas define them.

    enhance.py:172-181: S_k = (I - ((D^/alph)^T . D^))^T ; :183-195: W_k = D^/alph ; :197-204: b_k = -lam1/alph,
    with D^ = exp(log_D)/sqrt(sum(exp(log_D)^2, axis=0)).  S_k is None for k == 0 (custom_layers.py:363).
    """
    D = np.exp(np.asarray(p["log_D"][k], dtype=dtype))
    Dn = D / np.sqrt(np.sum(np.square(D), axis=0, keepdims=True))
    alph = np.exp(np.asarray(p["log_alph"][k], dtype=dtype))       # scalar or (R,), broadcasts over columns
    lam = np.exp(np.asarray(p["log_lam1"][k], dtype=dtype))
    R = D.shape[1]
    Wk = Dn / alph
    bk = -np.ones((R,), dtype=dtype) * lam / alph
    Sk = None
    if k > 0:
        Sk = (np.eye(R, dtype=dtype) - np.dot((Dn / alph).T, Dn)).T
    return Wk.astype(dtype), Sk, bk.astype(dtype)


def structured_U(p, dtype=np.float64):
    """(d0, o0, dk, ok): diagonal / off-diagonal value of U_0 = exp(log_U1)^T and U_k = exp(log_Uk)^T.

    enhance.py:163-167.  Raises if the matrices are not of the 'constant diagonal + constant off-diagonal'
    shape build_alt creates (the only shape any shipped config uses; they are non-trainable there).
    """
    out = []
    for key in ("log_U1", "log_Uk"):
        U = np.exp(np.asarray(p[key], dtype=dtype))
        R = U.shape[0]
        d = np.diag(U)
        off = U[~np.eye(R, dtype=bool)]
        if not (np.all(d == d[0]) and (off.size == 0 or np.all(off == off[0]))):
            raise ValueError("%s is not (a*I + b*11^T)-structured" % key)
        out += [dtype(d[0]), dtype(off[0] if off.size else 0.0)]
    return tuple(out)


# --------------------------------------------------------------------------------------------
# A.1  recurrence (custom_layers.py:336-375 step / get_initial_state; Keras masked K.rnn [unpinned])
# --------------------------------------------------------------------------------------------
def rnn_forward(x, p, mask_value=-1.0, dtype=np.float64, dense_U=False, return_all_hidden=False):
    """SimpleDeepRNN forward over a padded batch.  x: (B,T,F).  Returns H (B,T,R) [or (B,T,K*R)].

    custom_layers.py:343-375 (step), :336-341 (h0 tiled over the batch), :202-206 (h0 = softplus(log_h0)).
    Keras 2.0.4 semantics [unpinned]: Masking -> m_t = any_f(x != mask_value), x~ = x*m ; masked scan keeps
    the previous output (zeros before the first step) and the previous state where m_t is false.
    dense_U=True multiplies by the full R x R U_k like the reference does (custom_layers.py:362); the default
    uses the identical-in-exact-arithmetic structured form  prev*(d-o) + o*sum(prev).
    """
    x = np.asarray(x)
    B, T, F = x.shape
    K = int(np.asarray(p["log_D"]).shape[0])
    R = int(np.asarray(p["log_D"]).shape[2])
    m = np.any(x != mask_value, axis=-1)                                   # (B,T)
    xm = (x * m[..., None]).astype(dtype)
    Wk, Sk, bk = [], [], []
    for k in range(K):
        w, s, b = layer_weights(p, k, dtype)
        Wk.append(w), Sk.append(s), bk.append(b)
    if dense_U:
        U0 = np.exp(np.asarray(p["log_U1"], dtype=dtype)).T
        Uk = np.exp(np.asarray(p["log_Uk"], dtype=dtype)).T
    else:
        d0, o0, dk, ok = structured_U(p, dtype)
    lh0 = np.asarray(p["log_h0"], dtype=dtype)
    h0 = np.logaddexp(lh0, 0.0).astype(dtype)                              # softplus
    state = np.tile(h0[None, :], (B, 1))
    if return_all_hidden:
        raise NotImplementedError("flag_return_all_hidden is off the shipped path (SURVEY f4)")
    out_prev = np.zeros((B, R), dtype=dtype)
    H = np.zeros((B, T, R), dtype=dtype)
    # the input projections do not depend on the state: hoist them (same arithmetic, one GEMM per layer)
    XW = [np.matmul(xm.reshape(B * T, F), Wk[k]).reshape(B, T, R) for k in range(K)]
    for t in range(T):
        prev = state
        if not dense_U:
            psum = prev.sum(axis=1, keepdims=True)
        g = None
        for k in range(K):
            if dense_U:
                pre = np.dot(prev, U0 if k == 0 else Uk)
            else:
                d, o = (d0, o0) if k == 0 else (dk, ok)
                pre = prev * (d - o) + o * psum
            if k > 0:
                pre = pre + np.dot(g, Sk[k])
            pre = pre + XW[k][:, t, :]
            g = np.maximum(pre + bk[k], 0.0)
        mt = m[:, t][:, None]
        out_prev = np.where(mt, g, out_prev)
        state = np.where(mt, g, state)
        H[:, t, :] = out_prev
    return H


def output_head(H, p, transform_before_irm=None, dtype=np.float64):
    """enhance.py:269-305 + custom_layers.py:23-29 (DenseNonNegW) + :41-45 (DivideAbyAplusB)."""
    H = np.asarray(H, dtype=dtype)
    R = H.shape[-1]
    r = R // 2
    S = np.matmul(H[..., :r], np.exp(np.asarray(p["k_clean"], dtype=dtype)))
    N = np.matmul(H[..., r:], np.exp(np.asarray(p["k_noise"], dtype=dtype)))
    if transform_before_irm == "square":
        S, N = np.square(S), np.square(N)
    elif transform_before_irm is not None:
        raise ValueError("Unknown 'transform_before_irm' of '%s'" % transform_before_irm)
    e = dtype(EPS)
    return np.exp(np.log(e + S) - np.log(e + S + N))


def drnmf_forward(x, p, mask_value=-1.0, dtype=np.float64, transform_before_irm=None, dense_U=False):
    """Full model of build_unfolded_snmf: returns (H, irm)."""
    H = rnn_forward(x, p, mask_value, dtype, dense_U)
    return H, output_head(H, p, transform_before_irm, dtype)


def training_loss(x, y, irm, mask):
    """enhance.py:1040-1073: y^ = x_raw * irm, 'mse' over F, temporal sample weights = frame mask.

    Build-defined normalisation (Keras's weighted-mean rescaling is a per-batch constant [unpinned]):
        L = sum_{b,t} m * mean_F (x*irm - y)^2 / sum m
    """
    mask = np.asarray(mask, dtype=irm.dtype).reshape(irm.shape[0], irm.shape[1])
    per = np.mean(np.square(np.asarray(x, irm.dtype) * irm - np.asarray(y, irm.dtype)), axis=-1)
    return float(np.sum(per * mask) / np.sum(mask))


def param_count_notebook(F, r, K):
    """plot_learning_curves_waspaa2017.ipynb:121-126: K*F*2r + K + 2r  (log_D_k, log_alph_k, log_h0)."""
    return K * F * 2 * r + K + 2 * r


# --------------------------------------------------------------------------------------------
# A.2  frame-parallel ISTA (enhance.py:402-418, dead code in the reference; oracle "A")
# --------------------------------------------------------------------------------------------
def ista_ed(x, W, H, lam1, alph, K):
    xest = np.dot(W, H)
    for _ in range(K):
        H = np.maximum(0, -lam1 / alph + H + (1.0 / alph) * np.dot(W.T, x - xest))
        xest = np.dot(W, H)
    return H


# --------------------------------------------------------------------------------------------
# A.3  sparse NMF, Euclidean branch (sparseNMF/sparse_nmf_gpu.m) + chunk driver (snmf.py)
# --------------------------------------------------------------------------------------------
def sparse_nmf_ed(V, params, dtype=np.float64, rng=None, beta=2.0):
    """sparse_nmf_gpu.m:72-304.  beta = 2 is cf='ed' (every shipped config, enhance.py:568,590); beta = 1 ('kl', the
    solver's own default, :100-115) and beta = 0 ('is') and any other beta follow the generic branches (:212-226,
    :232-260, :266-276) - see sparse_nmf_beta below.

    params keys as in the .m: sparsity, max_iter, conv_eps, r, init_w, init_h ('ones' or array), w_update_ind,
    h_update_ind.  MATLAB's legacy rand('seed') stream (:119,:126,:132,:139) cannot be reproduced: missing
    initialisers are drawn from `rng` (numpy) instead -> parity only with explicit init_w/init_h.
    Returns (w, h, {'div': array, 'cost': array}).
    """
    V = np.asarray(V, dtype=dtype)
    m, n = V.shape
    max_iter = int(params.get("max_iter", 100))
    conv_eps = float(params.get("conv_eps", 0.0))
    if rng is None:
        rng = np.random.default_rng(int(params.get("random_seed", 1)))
    if "init_w" not in params or params["init_w"] is None:
        r = int(params["r"])
        w = rng.random((m, r)).astype(dtype)
    else:
        w = np.array(params["init_w"], dtype=dtype)
        ri = w.shape[1]
        if "r" in params and ri < int(params["r"]):                       # :129-133
            w = np.concatenate([w, rng.random((m, int(params["r"]) - ri)).astype(dtype)], axis=1)
        r = w.shape[1]
    ih = params.get("init_h", None)
    if ih is None:
        h = rng.random((r, n)).astype(dtype)
    elif isinstance(ih, str) and ih == "ones":
        h = np.ones((r, n), dtype=dtype)
    else:
        h = np.array(ih, dtype=dtype)
    w_ind = np.asarray(params.get("w_update_ind", np.ones(r, bool))).astype(bool).ravel()
    h_ind = np.asarray(params.get("h_update_ind", np.ones(r, bool))).astype(bool).ravel()
    sp = np.asarray(params.get("sparsity", 0.0), dtype=dtype)
    if sp.ndim == 0 or sp.size == 1:
        sp = np.ones((r, n), dtype=dtype) * sp.reshape(())                # :157-161
    elif sp.ndim == 1 or sp.shape[1] == 1:
        sp = np.repeat(sp.reshape(r, 1), n, axis=1)
    wn = np.sqrt(np.sum(w ** 2, axis=0))                                  # :164-166
    w = w / wn
    h = h * wn[:, None]
    flr = dtype(1e-9)
    lam = np.maximum(w @ h, flr)
    if beta != 2.0 and np.any(V == 0):                                    # :201-205
        V = V.copy()
        V[V == 0] = V[V > 0].min()
    last_cost = np.inf
    divs, costs = [], []
    update_h, update_w = h_ind.sum() > 0, w_ind.sum() > 0
    for it in range(1, max_iter + 1):
        if update_h:                                                      # :217-221, :228
            if beta == 2.0:
                P, Q = lam, V
            elif beta == 1.0:                                             # :212-216 (W^T 1 = column sums of W)
                P, Q = np.ones_like(lam), V / lam
            else:                                                         # :222-226
                P, Q = lam ** (beta - 1), V * lam ** (beta - 2)
            dph = w[:, h_ind].T @ P + sp[h_ind]
            dph = np.maximum(dph, flr)
            dmh = w[:, h_ind].T @ Q
            h[h_ind] = h[h_ind] * dmh / dph
            lam = np.maximum(w @ h, flr)
        if update_w:                                                      # :243-249, :262-263
            hw = h[w_ind]
            ww = w[:, w_ind]
            if beta == 2.0:
                P, Q = lam, V
            elif beta == 1.0:                                             # :232-241 (1 H^T = row sums of H)
                P, Q = np.ones_like(lam), V / lam
            else:                                                         # :250-259
                P, Q = lam ** (beta - 1), V * lam ** (beta - 2)
            VH = Q @ hw.T
            LH = P @ hw.T
            dpw = LH + np.sum(VH * ww, axis=0, keepdims=True) * ww
            dpw = np.maximum(dpw, flr)
            dmw = VH + np.sum(LH * ww, axis=0, keepdims=True) * ww
            w[:, w_ind] = ww * dmw / dpw
            w = w / np.sqrt(np.sum(w ** 2, axis=0))
            lam = np.maximum(w @ h, flr)
        if beta == 2.0:
            div = np.sum((V - lam) ** 2)                                  # :271  (no 1/2)
        elif beta == 1.0:
            div = np.sum(V * np.log(V / lam) - V + lam)                   # :269
        elif beta == 0.0:
            div = np.sum(V / lam - np.log(V / lam) - 1)                   # :273
        else:
            div = np.sum(V ** beta + (beta - 1) * lam ** beta - beta * V * lam ** (beta - 1)) / (beta * (beta - 1))
        cost = div + np.sum(sp * h)                                       # :278
        divs.append(float(div)), costs.append(float(cost))
        if it > 1 and conv_eps > 0:                                       # :288-296
            e = abs(cost - last_cost) / last_cost
            if e < conv_eps:
                break
        last_cost = cost
    return w, h, {"div": np.array(divs), "cost": np.array(costs)}


def sparse_nmf_beta(V, params, dtype=np.float64, rng=None):
    """sparse_nmf_gpu.m:100-115: cf in {'is', 'kl', 'ed'} selects beta = 0 / 1 / 2, otherwise params['beta'] (default 1)."""
    cf = params.get("cf", "kl")
    beta = {"is": 0.0, "kl": 1.0, "ed": 2.0}.get(cf, float(params.get("beta", 1.0)))
    return sparse_nmf_ed(V, params, dtype=dtype, rng=rng, beta=beta)


def sparse_nmf_chunked(V, params, frame_batch_size=None, save_H=True, dtype=np.float64, rng=None):
    """snmf.py:9-85 sparse_nmf_matlab: sequential chunks carrying W forward.

    frame_batch_size defaults to the reference rule 700000*200/r (snmf.py:33-35).  Extension over the
    reference: an array init_h is sliced per chunk (the reference would hand MATLAB a mis-sized matrix).
    """
    params_copy = copy.deepcopy(dict(params))
    n_feats, n_frames = V.shape
    r = int(params["r"])
    if frame_batch_size is None:
        frame_batch_size = int(float(700000) * (200.0 / float(r)))
    n_chunks = int(np.ceil(float(n_frames) / float(frame_batch_size)))
    H = np.zeros((r, n_frames), dtype=dtype) if save_H else None
    obj = {"obj_snmf_per_chunk": []}
    ic = fc = idv = fdv = 0.0
    init_h_full = params_copy.get("init_h", None)
    W = None
    for i in range(n_chunks):
        s, e = i * frame_batch_size, (i + 1) * frame_batch_size
        pc = dict(params_copy)
        if isinstance(init_h_full, np.ndarray):
            pc["init_h"] = init_h_full[:, s:e]
        W, H_tmp, o = sparse_nmf_ed(V[:, s:e], pc, dtype=dtype, rng=rng)
        if "w_update_ind" in params_copy:                                 # snmf.py:60-64
            idx = np.where(np.asarray(params_copy["w_update_ind"]).astype(bool))[0]
            params_copy["init_w"] = np.array(params_copy["init_w"], dtype=dtype)
            params_copy["init_w"][:, idx] = W[:, idx]
        else:
            params_copy["init_w"] = W
        obj["obj_snmf_per_chunk"].append(o)
        ic += o["cost"][0]; idv += o["div"][0]; fc += o["cost"][-1]; fdv += o["div"][-1]
        if save_H:
            H[:, s:e] = H_tmp
    obj["cost"] = [ic, fc]
    obj["div"] = [idv, fdv]
    if n_chunks == 1:
        obj = obj["obj_snmf_per_chunk"][0]
    return W, H, obj


def train_snmf(clean_frames, noisy_frames, params_snmf, noise_init, init_h_clean=None, init_h_noisy=None,
               init_w_clean=None, dtype=np.float64):
    """enhance.py:81-135 two-stage dictionary learning (caching stripped).

    noise_init replaces np.random.rand(*W.shape) drawn from the global seed-7654 stream (:110, enhance.py:7).
    """
    p1 = dict(params_snmf)
    if init_w_clean is not None:
        p1["init_w"] = init_w_clean
    if init_h_clean is not None:
        p1["init_h"] = init_h_clean
    W, H, obj = sparse_nmf_chunked(clean_frames, p1, dtype=dtype)
    r = int(params_snmf["r"])
    W_init = np.concatenate((W, np.asarray(noise_init, dtype=W.dtype)), axis=1)
    idx_update = np.concatenate((np.zeros(r, dtype=bool), np.ones(r, dtype=bool)))
    p2 = dict(params_snmf)
    p2.update({"r": 2 * r, "init_w": W_init, "w_update_ind": idx_update})
    if init_h_noisy is not None:
        p2["init_h"] = init_h_noisy
    Wn, Hn, objn = sparse_nmf_chunked(noisy_frames, p2, dtype=dtype)
    return Wn, Hn, objn


def snmf_irm(W_noisy, H, r):
    """enhance.py:848-852: irm = S^/(1e-9 + S^ + N^)."""
    clean_est = np.dot(W_noisy[:, :r], H[:r])
    noise_est = np.dot(W_noisy[:, r:], H[r:])
    return clean_est / (1e-9 + clean_est + noise_est)


# --------------------------------------------------------------------------------------------
# A.4  STFT / mask / iSTFT (util.py, audio_dataset.py)
# --------------------------------------------------------------------------------------------
def sqrt_hann(N):
    """audio_dataset.py:194: sqrt(scipy.signal.hann(N, sym=False).astype(float32)) -> float32."""
    n = np.arange(N, dtype=np.float64)
    hann = (0.5 - 0.5 * np.cos(2.0 * np.pi * n / N)).astype(np.float32)
    return np.sqrt(hann)


def stft_mc(x, N=1024, hop=None, window=None, dtype=np.complex64):
    """util.py:171-201 + librosa 0.5.1 core.stft(center=False) [unpinned: conj convention].

    x (nsampl,) or (nch,nsampl) -> X (N/2+1, nfram, nch).  Zero-pad to a hop multiple (:184-187), N zeros both
    ends (:189-190), frames at stride hop, window, FFT, first N/2+1 bins, conjugated.
    """
    if hop is None:
        hop = N // 2
    x = np.asarray(x)
    if x.ndim == 1:
        x = x.reshape(1, -1)
    nch, nsampl = x.shape
    nfram = int(math.ceil(float(nsampl) / float(hop)))
    npad = nfram * hop - nsampl
    x = np.concatenate((x, np.zeros((nch, npad), x.dtype)), axis=1)
    pad = np.zeros((nch, N), x.dtype)
    x = np.concatenate((pad, x, pad), axis=1)
    if window is None:
        window = np.ones(N, dtype=x.dtype)       # librosa default is hann; the reference always passes one
    real_t = np.float32 if np.dtype(dtype) == np.complex64 else np.float64
    w = np.asarray(window, dtype=real_t).reshape(-1, 1)
    nT = 1 + (x.shape[1] - N) // hop
    X = np.zeros((N // 2 + 1, nT, nch), dtype=dtype)
    idx = np.arange(N)[:, None] + hop * np.arange(nT)[None, :]
    for ich in range(nch):
        frames = x[ich].astype(real_t)[idx]                                # (N, nT)
        X[:, :, ich] = np.fft.fft(w * frames, axis=0)[: N // 2 + 1].conj().astype(dtype)
    return X


def stack_reim(X):
    """util.py:351: Yaug = [Re(Y); Im(Y)] of shape (2F, frames) (single channel)."""
    Y = X[:, :, 0] if X.ndim == 3 else X
    return np.concatenate((np.real(Y), np.imag(Y)), axis=0)


def magnitude(Yaug):
    """audio_dataset.py:22-23 transform 'mag'."""
    F = Yaug.shape[0] // 2
    return np.sqrt(Yaug[:F] ** 2 + Yaug[F:] ** 2)


def istft_no_div(stft_matrix, hop_length, window, dtype=np.float32):
    """util.py:48-169 with center=False: window*(2.0/(N//hop)) (:143, py2 integer division), overlap-add of
    Re(ifft([conj(S), S[-2:0:-1]])) (:151-157), no window-sum division (:158-164)."""
    n_fft = 2 * (stft_matrix.shape[0] - 1)
    win = np.asarray(window)
    ifft_window = win * (2.0 / (n_fft // hop_length))
    n_frames = stft_matrix.shape[1]
    y = np.zeros(n_fft + hop_length * (n_frames - 1), dtype=dtype)
    for i in range(n_frames):
        spec = stft_matrix[:, i].flatten()
        spec = np.concatenate((spec.conj(), spec[-2:0:-1]), 0)
        ytmp = ifft_window * np.fft.ifft(spec).real
        y[i * hop_length: i * hop_length + n_fft] += ytmp.astype(dtype)
    return y


def istft_mc(X, hop, window, dtype=np.float32, nsampl=None):
    """util.py:203-226 with flag_noDiv=1: per channel istft_noDiv, then drop the last N and the first N samples."""
    N = 2 * (X.shape[0] - 1)
    nch = X.shape[2]
    rows = [istft_no_div(X[:, :, ich], hop, window, dtype) for ich in range(nch)]
    xr = np.stack(rows, axis=0)
    xr = xr[:, 0:(xr.shape[1] - N)]
    xr = xr[:, N:]
    if nsampl is not None:
        xr = xr[:, 0:nsampl]
    return xr, N


def reconstruct_x(x_stack, hop, window, mask=None, dtype=np.float32):
    """audio_dataset.py:267-278: tile the real mask over [Re;Im], multiply, complex, istft_mc(flag_noDiv=1)."""
    X = np.asarray(x_stack)
    if mask is not None:
        if mask.shape[0] < X.shape[0]:
            mask = np.tile(mask, (X.shape[0] // mask.shape[0], 1))
        X = mask * X
    F = X.shape[0] // 2
    Xc = (X[:F] + 1j * X[F:])[:, :, None]
    xr, _ = istft_mc(Xc, hop, window, dtype)
    return xr


def wav_quantize(x):
    """util.py:37-45 wavwrite + :29-35 wavread round trip of a float signal."""
    x = np.asarray(x, dtype=np.float32)
    mx = np.max(np.abs(x))
    if mx > 1:
        x = x / mx
    return np.int16(x * 32767.0).astype(np.float32) / 32768.0


def sdr_db(est, ref):
    """Single-source BSS-Eval SDR as called at score_audio.m:199-206 [unpinned: toolbox not vendored]:
    s_t = (<est,ref>/|ref|^2) ref ; SDR = 10 log10(|s_t|^2 / |est - s_t|^2), both cut to the shorter length."""
    est = np.asarray(est, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    n = min(est.size, ref.size)
    est, ref = est[:n], ref[:n]
    st = (np.dot(est, ref) / np.dot(ref, ref)) * ref
    return 10.0 * np.log10(np.dot(st, st) / np.dot(est - st, est - st))


def masked_seqs_to_frames(x, mask):
    """util.py:19-27."""
    n_examples, time_steps, n_feature = x.shape
    x = x.transpose((2, 0, 1))
    x_reshape = np.reshape(x, (n_feature, n_examples * time_steps))
    mask = mask.transpose((2, 0, 1))
    mask_reshape = np.reshape(mask, (n_examples * time_steps,))
    idx = np.where(mask_reshape == mask_reshape[0])[0]
    return x_reshape[:, idx]


# --------------------------------------------------------------------------------------------
# A.6  data formats either side of the path (audio_dataset.py, enhance.py:74-79)
# --------------------------------------------------------------------------------------------
def get_mask_value(config):
    """audio_dataset.py:11-17: -1 pads magnitude (or log-magnitude-target) data, 0 otherwise."""
    if config.get("transform_x") == "mag":
        return -1.0
    if config.get("transform_y") == "logmag":
        return -1.0
    return 0.0


def data_transform(kind):
    """audio_dataset.py:22-37: feature maps applied to a [Re;Im] stack (2F, frames) -> (F, frames)."""
    if kind == "mag":
        return lambda x: np.sqrt(x[:x.shape[0] // 2, :] ** 2 + x[x.shape[0] // 2:, :] ** 2)
    if kind == "logmag":
        return lambda x: np.log(np.float32(1.0) + np.sqrt(x[:x.shape[0] // 2, :] ** 2 + x[x.shape[0] // 2:, :] ** 2))
    return lambda x: x


def reshape_and_pad_stacks(x_stack, y_stack, fidx, transform_x=(lambda x: x), transform_y=(lambda y: y), pad_value=0.0,
                           maxlen=None):
    """audio_dataset.py:116-169: (2F, total frames) stacks -> (n_sequences, maxlen, d) padded with pad_value, plus a
    (n_sequences, maxlen, 1) mask.  With maxlen < longest file, files are cut into consecutive chunks of maxlen
    frames (:126-132, :149-167); chunks never span two files."""
    fidx = np.asarray(fidx)
    maxseq = int(np.max(fidx[:, 1] - fidx[:, 0]))
    if maxlen is None or maxlen > maxseq:
        maxlen = maxseq
    maxlen = int(maxlen)
    d = transform_x(x_stack[:, 0:1]).shape[0]
    if maxlen == maxseq:
        n_sequences = fidx.shape[0]
    else:
        n_sequences = 0
        for i in range(fidx.shape[0]):
            t = 0
            while t < (fidx[i, 1] - fidx[i, 0]):
                n_sequences += 1
                t += maxlen
    x = (pad_value * np.ones((n_sequences, maxlen, d))).astype(x_stack.dtype)
    y = (pad_value * np.ones((n_sequences, maxlen, d))).astype(y_stack.dtype)
    mask = np.zeros((n_sequences, maxlen, 1)).astype(x_stack.dtype)
    t = 0
    i_wavfile = 0
    for i in range(n_sequences):
        t_end = t + maxlen
        increment = False
        if t_end >= fidx[i_wavfile, 1]:
            t_end = int(fidx[i_wavfile, 1])
            increment = True
        x[i, :t_end - t, :] = transform_x(x_stack[:, t:t_end]).T
        y[i, :t_end - t, :] = transform_y(y_stack[:, t:t_end]).T
        mask[i, :t_end - t, :] = 1.0
        if increment and i < n_sequences - 1:
            i_wavfile += 1
            t = int(fidx[i_wavfile, 0])
        else:
            t += maxlen
    return x, y, mask


def clip_x_to_y(x, y, xfidx, yfidx):
    """audio_dataset.py:90-104: per utterance keep the first len(y_utt) frames of x (in place, then truncated)."""
    ylens = yfidx[:, 1] - yfidx[:, 0]
    idx = 0
    for iutt in range(xfidx.shape[0]):
        xcur = x[:, xfidx[iutt, 0]:xfidx[iutt, 1]]
        x[:, idx:idx + ylens[iutt]] = xcur[:, 0:ylens[iutt]]
        idx += ylens[iutt]
    return x[:, 0:y.shape[1]]


def snr_db(est, ref):
    """score_audio.m:209: raw SNR = 10 log10(sum(ref^2) / sum((ref - est)^2)) after truncation to the shorter (:199-204)."""
    n = min(len(est), len(ref))
    est, ref = np.asarray(est[:n], np.float64), np.asarray(ref[:n], np.float64)
    return 10.0 * np.log10(np.sum(ref ** 2) / np.sum((ref - est) ** 2))


def snmf_savefile_stem(params_snmf, path_dicts=""):
    """enhance.py:74-79 get_snmf_savefile without the extension: path_dicts + 'W_noisy_' + md5(json.dumps(params,
    sort_keys=True, cls=MyEncoder)) + '_sparsity%.3f'  (MyEncoder, :60-71, turns numpy scalars/arrays into python)."""
    import hashlib
    import json

    class _Enc(json.JSONEncoder):
        def default(self, obj):
            if isinstance(obj, np.integer):
                return int(obj)
            if isinstance(obj, np.floating):
                return float(obj)
            if isinstance(obj, np.ndarray):
                return obj.tolist()
            return super().default(obj)
    h = hashlib.md5(json.dumps(params_snmf, sort_keys=True, cls=_Enc).encode()).hexdigest()
    return path_dicts + "W_noisy_" + h + ("_sparsity%.3f" % params_snmf["sparsity"])
