"""Thin torch-tensor wrapper of the libdrnmf C-ABI.

PyTorch is plumbing here: device memory, the current CUDA stream and (for training) torch.distributed.  Every
number is produced by the hand-written sm_100a kernels behind include/drnmf.h; nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

PARAM_KEYS = ("log_D", "log_alph", "log_lam1", "log_U1", "log_Uk", "log_h0", "k_clean", "k_noise")


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.DrnmfError(5, "no CUDA device visible: the DR-NMF path runs on B200 only (no CPU fallback)")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def structured_u(log_U, what="log_U"):
    """(diag, off) of exp(log_U)^T, verifying the a*I + b*11^T structure build_alt creates (enhance.py:163-167)."""
    U = torch.as_tensor(np.asarray(log_U) if not torch.is_tensor(log_U) else log_U).detach().float().cpu()
    R = U.shape[0]
    if U.ndim != 2 or U.shape[1] != R:
        raise ValueError("%s must be square" % what)
    d = torch.diagonal(U)
    eye = torch.eye(R, dtype=torch.bool)
    off = U[~eye]
    if not (bool((d == d[0]).all()) and (off.numel() == 0 or bool((off == off[0]).all()))):
        raise NotImplementedError(
            "%s is not of the (a*I + b*11^T) form build_alt creates; a dense recurrent U is not supported by the "
            "fixed kernel (custom_layers.py:362 multiplies by it densely)" % what)
    dv = float(torch.exp(d[0]))
    ov = float(torch.exp(off[0])) if off.numel() else 0.0
    return dv, ov


class DrnmfEngine:
    """One DR-NMF model instance on the current CUDA device (drnmf_create .. drnmf_destroy)."""

    def __init__(self, F, R, K_layers, square_irm=False, impl=None):
        _require_cuda()
        self.lib = _lib.load()
        self.F, self.R, self.K = int(F), int(R), int(K_layers)
        flags = (_lib.FLAG_SQUARE_IRM if square_irm else 0)
        if impl == "simt":
            flags |= _lib.IMPL_SIMT
        h = C.c_void_p()
        _lib.check(self.lib.drnmf_create(C.byref(h), self.F, self.R, self.K, flags))
        self.h = h
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._ws = None
        self._ews = None
        self._tws = None
        self._keep = None
        rp, fp = C.c_int(), C.c_int()
        _lib.check(self.lib.drnmf_padded_dims(self.h, C.byref(rp), C.byref(fp)))
        self.Rp, self.Fp = rp.value, fp.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.drnmf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters ---------------------------------------------------------------------------
    def set_params(self, p):
        """p: dict with PARAM_KEYS (numpy or torch).  log_D (K|1,F,R) or (F,R); log_alph (K|1,) / (K|1,R) / scalar;
        log_lam1 (K|1,) or scalar; log_U1/log_Uk (R,R) or (diag, off) tuples; log_h0 (R,); k_clean/k_noise (R/2,F)."""
        dev = self.device

        def t(a):
            if torch.is_tensor(a):
                return a.detach().to(dev, torch.float32).contiguous()
            return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float32)), device=dev)

        log_D = t(p["log_D"])
        if log_D.ndim == 2:
            log_D = log_D[None]
        log_alph = t(p["log_alph"])
        if log_alph.ndim == 0:
            log_alph = log_alph.reshape(1, 1)
        elif log_alph.ndim == 1:
            # (K,) scalars per layer, or (R,) untied-alph vector of a tied model
            if log_alph.shape[0] in (1, self.K):
                log_alph = log_alph.reshape(-1, 1)
            elif log_alph.shape[0] == self.R:
                log_alph = log_alph.reshape(1, -1)
            else:
                raise ValueError("log_alph has %d entries; expected 1, K or R" % log_alph.shape[0])
        log_lam1 = t(p["log_lam1"]).reshape(-1)
        log_h0 = t(p["log_h0"]).reshape(-1)
        k_clean, k_noise = t(p["k_clean"]), t(p["k_noise"])
        if tuple(log_D.shape[1:]) != (self.F, self.R):
            raise ValueError("log_D must be (K, F, R) = (*, %d, %d), got %s" % (self.F, self.R, tuple(log_D.shape)))
        if tuple(k_clean.shape) != (self.R // 2, self.F) or tuple(k_noise.shape) != (self.R // 2, self.F):
            raise ValueError("k_clean / k_noise must be (R/2, F) Keras Dense kernels")
        if log_h0.numel() != self.R:
            raise ValueError("log_h0 must have R entries")
        u1 = p["log_U1"] if isinstance(p["log_U1"], tuple) else structured_u(p["log_U1"], "log_U1")
        uk = p["log_Uk"] if isinstance(p["log_Uk"], tuple) else structured_u(p["log_Uk"], "log_Uk")
        self._keep = (log_D, log_alph, log_lam1, log_h0, k_clean, k_noise)
        _lib.check(self.lib.drnmf_set_params(
            self.h, _ptr(log_D), log_D.shape[0], _ptr(log_alph), log_alph.shape[0], log_alph.shape[1], _ptr(log_lam1),
            log_lam1.shape[0], _ptr(log_h0), _ptr(k_clean), _ptr(k_noise), u1[0], u1[1], uk[0], uk[1], _stream()))

    # ---- inference ----------------------------------------------------------------------------
    def _workspace(self, nbytes, which="_ws"):
        ws = getattr(self, which)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=self.device)
            setattr(self, which, ws)
        off = (-ws.data_ptr()) % 256
        return C.c_void_p(ws.data_ptr() + off), ws.numel() - off

    def forward(self, x, mask_value=-1.0, want_H=True, want_irm=True, H_out=None, irm_out=None):
        """x: (B,T,F) float32 CUDA tensor padded with mask_value -> (H (B,T,R) | None, irm (B,T,F) | None)."""
        if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32):
            raise TypeError("x must be a float32 CUDA tensor (the product path has no CPU implementation)")
        x = x.contiguous()
        B, T, F = x.shape
        if F != self.F:
            raise ValueError("x has %d features, model expects %d" % (F, self.F))
        H = H_out if H_out is not None else (
            torch.empty((B, T, self.R), dtype=torch.float32, device=x.device) if want_H else None)
        irm = irm_out if irm_out is not None else (
            torch.empty((B, T, F), dtype=torch.float32, device=x.device) if want_irm else None)
        need = self.lib.drnmf_workspace_bytes(self.h, B, T)
        ws, wsb = self._workspace(need)
        _lib.check(self.lib.drnmf_forward(self.h, _ptr(x), B, T, float(mask_value), _ptr(H), _ptr(irm), ws, wsb,
                                          _stream()))
        return H, irm

    def enhance_host(self, x_host, stack_host, frames_host, N, hop, mask_value=-1.0, out=None):
        """End-to-end with HOST tensors (pinned recommended): magnitudes (B,T,F) + [Re;Im] stack (2F, B*T) ->
        enhanced audio (B, hop*(T-1)-N) on the host.  H2D/D2H copies happen inside the call."""
        B, T, F = x_host.shape
        L = hop * (T - 1) - N
        if out is None:
            out = torch.empty((B, L), dtype=torch.float32).pin_memory()
        need = self.lib.drnmf_enhance_workspace_bytes(self.h, B, T, N, hop)
        ws, wsb = self._workspace(need, "_ews")
        _lib.check(self.lib.drnmf_enhance_host(self.h, _ptr(x_host), _ptr(stack_host), _ptr(frames_host), B, T, N, hop,
                                               float(mask_value), _ptr(out), ws, wsb, _stream()))
        return out

    def loss_and_grads(self, x, y, mask_value=-1.0, want_irm=False, out=None, layer_ready=None):
        """Training step on the device: forward with stored activations, masked-MSE loss (enhance.py:1040-1073) and
        hand-written BPTT.  Returns (loss_sum, mask_sum, grads dict of CUDA tensors [, irm]); the loss is
        loss_sum / mask_sum and the gradients are those of loss_sum (divide by the all-reduced frame count).
        out: dict of preallocated gradient tensors (same keys / shapes as returned) to write into.
        layer_ready(k): called right after the kernels producing the gradients of layer k have been enqueued on the
        current stream (drnmf_loss_and_grads_cb) - the hook for per-layer gradient all-reduce under the later GEMMs."""
        if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32):
            raise TypeError("x must be a float32 CUDA tensor")
        x, y = x.contiguous(), y.contiguous()
        B, T, F = x.shape
        dev = x.device
        log_D, log_alph, log_lam1 = self._keep[0], self._keep[1], self._keep[2]
        if out is None:
            g = {"log_D": torch.zeros_like(log_D), "log_alph": torch.zeros_like(log_alph).reshape(-1),
                 "log_lam1": torch.zeros_like(log_lam1), "log_h0": torch.zeros(self.R, device=dev),
                 "k_clean": torch.zeros((self.R // 2, F), device=dev), "k_noise": torch.zeros((self.R // 2, F), device=dev)}
        else:
            g = out
            want = {"log_D": log_D.numel(), "log_alph": log_alph.numel(), "log_lam1": log_lam1.numel(), "log_h0": self.R,
                    "k_clean": (self.R // 2) * F, "k_noise": (self.R // 2) * F}
            for k, n in want.items():
                if not (g[k].is_cuda and g[k].dtype == torch.float32 and g[k].is_contiguous() and g[k].numel() == n):
                    raise ValueError("out[%r] must be a contiguous float32 CUDA tensor with %d elements" % (k, n))
        irm = torch.empty((B, T, F), dtype=torch.float32, device=dev) if want_irm else None
        need = self.lib.drnmf_train_workspace_bytes(self.h, B, T)
        ws, wsb = self._workspace(need, "_tws")
        loss = (C.c_double * 2)()
        failed = []

        def _cb(user, k, stream):
            try:
                layer_ready(int(k))
                return 0
            except Exception as e:      # an exception must not unwind through the C frames
                failed.append(e)
                return 1
        cb = _lib.LAYER_FN(_cb) if layer_ready is not None else _lib.LAYER_FN(0)
        rc = self.lib.drnmf_loss_and_grads_cb(self.h, _ptr(x), _ptr(y), B, T, float(mask_value), _ptr(g["log_D"]),
                                              _ptr(g["log_alph"]), _ptr(g["log_lam1"]), _ptr(g["log_h0"]),
                                              _ptr(g["k_clean"]), _ptr(g["k_noise"]), loss, _ptr(irm), ws, wsb, _stream(), cb, None)
        if failed:
            raise failed[0]
        _lib.check(rc)
        res = (float(loss[0]), float(loss[1]), g)
        return res + (irm,) if want_irm else res

    def set_training_loss(self, kind="mse_of_masked", lam1=0.0):
        """'mse_of_masked' (enhance.py:1040-1047) or 'snmf_cost' = the optional pretraining objective of enhance.py:1024-1036."""
        kinds = {"mse_of_masked": 0, "snmf_cost": 1}
        if kind not in kinds:
            raise ValueError("Unknown 'loss' of '%s'" % kind)        # the reference constructs and drops this error
        _lib.check(self.lib.drnmf_set_training_loss(self.h, kinds[kind], float(lam1)))

    def forward_all_hidden(self, x, mask_value=-1.0):
        """flag_return_all_hidden (custom_layers.py:371-374): (B,T,K*R) hidden vectors of all layers, layer-major."""
        x = x.contiguous()
        B, T, F = x.shape
        out = torch.empty((B, T, self.K * self.R), dtype=torch.float32, device=x.device)
        need = self.lib.drnmf_train_workspace_bytes(self.h, B, T)
        ws, wsb = self._workspace(need, "_tws")
        _lib.check(self.lib.drnmf_forward_all_hidden(self.h, _ptr(x), B, T, float(mask_value), _ptr(out), ws, wsb, _stream()))
        return out

    def adam_step(self, params, grads, m, v, lr_t, beta_1=0.9, beta_2=0.999, epsilon=1e-8, grad_scale=1.0, trainable=None):
        """Fused Keras-formula Adam on flat float32 CUDA buffers (drnmf_adam_step); `trainable`: uint8 mask or None."""
        n = params.numel()
        for t in (params, grads, m, v):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n):
                raise ValueError("adam_step needs contiguous float32 CUDA buffers of equal size")
        if trainable is not None and not (trainable.is_cuda and trainable.dtype == torch.uint8 and trainable.numel() == n):
            raise ValueError("trainable must be a uint8 CUDA tensor with one byte per parameter")
        _lib.check(self.lib.drnmf_adam_step(_ptr(params), _ptr(grads), _ptr(m), _ptr(v), _ptr(trainable), n, float(lr_t),
                                            float(beta_1), float(beta_2), float(epsilon), float(grad_scale), _stream()))

    def stage_times(self):
        """ms of (masking, projection GEMM, recurrence, recon+mask GEMM) of the last forward (CUDA events)."""
        ms = (C.c_float * 4)()
        _lib.check(self.lib.drnmf_stage_times(self.h, ms))
        return [float(v) for v in ms]

    def recurrent_config(self, backward=False):
        """dict describing how the recurrence of the last forward (or the backward chain of the last loss_and_grads) ran:
        impl 'tcgen05' | 'simt' + tiling; n_tiles = batch tiles per group, groups = batch groups on disjoint SMs."""
        c = (C.c_int * 10)()
        _lib.check(self.lib.drnmf_recurrent_config2(self.h, 1 if backward else 0, c))
        keys = ("NB", "KS", "MT", "NSC", "n_tiles", "WST", "HST", "RST", "groups")
        d = {"impl": "tcgen05" if c[0] == 0 else "simt"}
        d.update({k: int(c[1 + i]) for i, k in enumerate(keys)})
        return d

    def last_backward_impl(self):
        return self.recurrent_config(backward=True)["impl"]

    def inject_device_error(self, code):
        """Test hook (drnmf_debug_inject_error): latch `code` in the device-side error word."""
        _lib.check(self.lib.drnmf_debug_inject_error(self.h, int(code), _stream()))

    def derived(self, which, k=0):
        n = {0: self.Rp * self.Rp, 1: self.Rp * self.Fp, 2: self.Rp, 3: self.Rp}[which]
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.drnmf_get_derived(self.h, which, k, _ptr(out), _stream()))
        return out.reshape({0: (self.Rp, self.Rp), 1: (self.Rp, self.Fp), 2: (self.Rp,), 3: (self.Rp,)}[which])


class EnhancePlan:
    """Preallocated device-resident enhancement step: forward -> mask -> iSTFT with inputs already in HBM
    (enhance.py:1186-1203 without the host copies).  frames[b] = valid frames of utterance b (default T)."""

    def __init__(self, eng, B, T, N, hop, frames=None):
        self.eng, self.B, self.T, self.N, self.hop = eng, B, T, N, hop
        dev = eng.device
        F = eng.F
        fr = np.full((B,), T, np.int64) if frames is None else np.asarray(frames, np.int64)
        self.L = hop * (T - 1) - N
        starts = np.arange(B, dtype=np.int64) * T
        self.fidx = torch.as_tensor(np.stack([starts, starts + fr], axis=1).copy(), device=dev)
        self.out_offs = torch.as_tensor(np.arange(B, dtype=np.int64) * self.L, device=dev)
        self.irm = torch.empty((B, T, F), dtype=torch.float32, device=dev)
        self.audio = torch.zeros((B, self.L), dtype=torch.float32, device=dev)
        self.nb = eng.lib.drnmf_istft_workspace_bytes(B * T, N)
        self.ws = torch.empty(self.nb + 256, dtype=torch.uint8, device=dev)
        self.ws_ptr = C.c_void_p(self.ws.data_ptr() + ((-self.ws.data_ptr()) % 256))
        self.max_frames = int(fr.max())

    def run(self, x, stack, mask_value=-1.0):
        eng = self.eng
        eng.forward(x, mask_value, want_H=False, irm_out=self.irm)
        _lib.check(eng.lib.drnmf_mask_istft(_ptr(stack), _ptr(self.irm), _ptr(self.fidx), _ptr(self.out_offs), self.B,
                                            self.max_frames, self.N, self.hop, self.B * self.T, _ptr(self.audio),
                                            self.ws_ptr, self.nb, _stream()))
        return self.audio


# ---- STFT / iSTFT ------------------------------------------------------------------------------
_STFT_TABLES = {}


def stft_frames(nsampl, N, hop):
    return _lib.load().drnmf_stft_frames(int(nsampl), int(N), int(hop))


def stft_mag(audio, offs, lens, N, hop, want_stack=True, want_mag=True):
    """audio: 1-D float32 CUDA tensor holding all utterances; offs/lens: per-utterance start/length (python lists).
    Returns (stack (2F, total) | None, mag (total, F) | None, fidx (n_utt,2) int64 CUDA tensor)."""
    _require_cuda()
    lib = _lib.load()
    dev = audio.device
    n_utt = len(lens)
    # the per-utterance index tables only depend on the layout of the batch: built (and uploaded) once per layout
    key = (str(dev), int(N), int(hop), tuple(int(n) for n in lens), tuple(int(o) for o in offs))
    hit = _STFT_TABLES.get(key)
    if hit is None:
        frames = [lib.drnmf_stft_frames(int(n), N, hop) for n in lens]
        starts = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
        hit = (int(starts[-1]), max(frames) if frames else 0,
               torch.as_tensor(np.stack([starts[:-1], starts[1:]], axis=1).copy(), device=dev),
               torch.as_tensor(np.asarray(offs, dtype=np.int64), device=dev),
               torch.as_tensor(np.asarray(lens, dtype=np.int32), device=dev))
        if len(_STFT_TABLES) >= 16:
            _STFT_TABLES.clear()
        _STFT_TABLES[key] = hit
    total, max_fr, fidx, offs_t, lens_t = hit
    frames = [max_fr]
    F = N // 2 + 1
    stack = torch.empty((2 * F, total), dtype=torch.float32, device=dev) if want_stack else None
    mag = torch.empty((total, F), dtype=torch.float32, device=dev) if want_mag else None
    _lib.check(lib.drnmf_stft_mag(_ptr(audio), _ptr(offs_t), _ptr(lens_t), _ptr(fidx), n_utt, max(frames) if frames else 0,
                                  N, hop, total, _ptr(stack), _ptr(mag), _stream()))
    return stack, mag, fidx


def mask_istft(stack, mask, fidx, N, hop):
    """stack (2F,total) CUDA, mask (total,F) CUDA or None, fidx (n_utt,2) int64 -> list of 1-D CUDA tensors."""
    _require_cuda()
    lib = _lib.load()
    dev = stack.device
    fi = fidx.detach().cpu().numpy().astype(np.int64)
    n_utt = fi.shape[0]
    counts = (fi[:, 1] - fi[:, 0]).astype(np.int64)
    out_lens = np.maximum(hop * (counts - 1) - N, 0)
    out_offs = np.concatenate([[0], np.cumsum(out_lens)]).astype(np.int64)
    out = torch.zeros(int(out_offs[-1]) + 1, dtype=torch.float32, device=dev)
    total = stack.shape[1]
    nb = lib.drnmf_istft_workspace_bytes(total, N)
    ws = torch.empty(nb + 256, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 256
    fidx_d = fidx.to(dev, torch.int64).contiguous()
    offs_d = torch.as_tensor(out_offs[:-1].copy(), device=dev)
    _lib.check(lib.drnmf_mask_istft(_ptr(stack), _ptr(mask), _ptr(fidx_d), _ptr(offs_d), n_utt,
                                    int(counts.max()) if n_utt else 0, N, hop, total, _ptr(out),
                                    C.c_void_p(ws.data_ptr() + off), nb, _stream()))
    return [out[int(out_offs[u]):int(out_offs[u + 1])] for u in range(n_utt)]


# ---- sparse NMF multiplicative updates -------------------------------------------------------------
def snmf_mu_ed(V, W, H, sparsity, max_iter, conv_eps=0.0, w_update=None, h_update=None, impl=None, group=None,
               distributed=False, beta=2.0):
    """V (F,n), W (F,R), H (R,n): float32 CUDA tensors (W and H are updated IN PLACE).  w_update / h_update: boolean
    arrays of length R (None = all).  Returns (cost, div) numpy arrays truncated at convergence
    (sparse_nmf_gpu.m:288-296).  Replaces the MATLAB subprocess of snmf.py:88-113.  beta selects the divergence
    (sparse_nmf_gpu.m:100-115): 2 = 'ed' (every shipped config), 1 = 'kl', 0 = 'is', else the generic branch."""
    _require_cuda()
    lib = _lib.load()
    for t in (V, W, H):
        if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise TypeError("V, W, H must be contiguous float32 CUDA tensors")
    F, n = V.shape
    R = W.shape[1]
    if tuple(W.shape) != (F, R) or tuple(H.shape) != (R, n):
        raise ValueError("shape mismatch: V %s W %s H %s" % (tuple(V.shape), tuple(W.shape), tuple(H.shape)))
    max_iter = int(max_iter)
    cost = np.zeros(max_iter, np.float64)
    div = np.zeros(max_iter, np.float64)
    iters = C.c_int(0)

    def mask(m):
        if m is None:
            return None, C.c_void_p(0)
        a = np.ascontiguousarray(np.asarray(m).astype(bool).ravel().astype(np.uint8))
        if a.size != R:
            raise ValueError("update mask must have R entries")
        return a, C.c_void_p(a.ctypes.data)

    wm, wp = mask(w_update)
    hm, hp = mask(h_update)
    beta = float(beta)
    nb = lib.drnmf_snmf_beta_workspace_bytes(F, n, R, beta)
    ws = torch.empty(nb + 256, dtype=torch.uint8, device=V.device)
    off = (-ws.data_ptr()) % 256
    flags = _lib.IMPL_SIMT if impl == "simt" else 0
    import torch.distributed as dist
    if distributed and not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("snmf_mu_ed(distributed=True) needs an initialised torch.distributed process group")
    cb = _lib.ALLREDUCE_FN(0)
    if distributed:
        # frames (columns of V, H) are sharded over ranks, W is replicated: the library asks for the sums of V H^T,
        # Lambda H^T and of the cost terms through this callback; the buffers live inside `ws`, so they are exposed to
        # torch.distributed as zero-copy views of it.
        base = ws.data_ptr()

        def _allreduce(user, ptr, count, dtype, stream):
            try:
                item, tdt = (8, torch.float64) if dtype == 1 else (4, torch.float32)
                o = int(ptr) - base
                view = ws[o:o + count * item].view(tdt)
                dist.all_reduce(view, op=dist.ReduceOp.MIN if dtype == 2 else dist.ReduceOp.SUM, group=group)
                return 0
            except Exception as e:      # an exception must not unwind through the C frames
                import sys
                print("[drnmf] all-reduce callback failed: %r" % (e,), file=sys.stderr)
                return 1
        cb = _lib.ALLREDUCE_FN(_allreduce)
    _lib.check(lib.drnmf_snmf_mu_beta(F, n, R, beta, _ptr(V), _ptr(W), _ptr(H), wp, hp, float(sparsity), max_iter,
                                      float(conv_eps), C.c_void_p(cost.ctypes.data), C.c_void_p(div.ctypes.data),
                                      C.byref(iters), flags, C.c_void_p(ws.data_ptr() + off), nb, _stream(), cb, None))
    k = iters.value
    return cost[:k].copy(), div[:k].copy()


def snmf_irm(W, H, r, impl=None):
    """enhance.py:847-852 on the GPU: W (F,R), H (R,n) float32 CUDA tensors -> irm (F,n) = S^/(1e-9 + S^ + N^)."""
    _require_cuda()
    lib = _lib.load()
    W, H = W.contiguous(), H.contiguous()
    F, R = W.shape
    n = H.shape[1]
    if H.shape[0] != R or n % 4 != 0:
        raise ValueError("H must be (R, n) with n a multiple of 4 (16-byte rows); pad with zero frames")
    out = torch.empty((F, n), dtype=torch.float32, device=W.device)
    nb = lib.drnmf_snmf_irm_workspace_bytes(F, n, R)
    ws = torch.empty(nb + 256, dtype=torch.uint8, device=W.device)
    off = (-ws.data_ptr()) % 256
    _lib.check(lib.drnmf_snmf_irm(F, n, R, int(r), _ptr(W), _ptr(H), _ptr(out), _lib.IMPL_SIMT if impl == "simt" else 0,
                                  C.c_void_p(ws.data_ptr() + off), nb, _stream()))
    return out


def ista_ed(x, W, H, lam1, alph, K, impl=None):
    """enhance.py:402-418 on the GPU: x (F,n), W (F,R), H (R,n) float32 CUDA tensors; returns the updated H (new tensor)."""
    _require_cuda()
    lib = _lib.load()
    F, n = x.shape
    R = W.shape[1]
    pad = (-n) % 4
    if pad:   # frames are independent: pad with zero frames, drop them afterwards
        x = torch.cat([x, x.new_zeros(F, pad)], dim=1)
        H = torch.cat([H, H.new_zeros(R, pad)], dim=1)
    x, W, Hn = x.contiguous(), W.contiguous(), H.contiguous().clone()
    nb = lib.drnmf_ista_workspace_bytes(F, n + pad, R)
    ws = torch.empty(nb + 256, dtype=torch.uint8, device=x.device)
    off = (-ws.data_ptr()) % 256
    _lib.check(lib.drnmf_ista_ed(F, n + pad, R, _ptr(x), _ptr(W), _ptr(Hn), float(lam1), float(alph), int(K),
                                 _lib.IMPL_SIMT if impl == "simt" else 0, C.c_void_p(ws.data_ptr() + off), nb, _stream()))
    return Hn[:, :n]
