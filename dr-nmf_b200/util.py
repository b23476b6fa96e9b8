"""Host-side mirror of the reference's util.py STFT helpers over the CUDA kernels (util.py:171-226).

stft_mc / istft_mc keep the reference's array conventions: x (nch, nsampl) or (nsampl,) -> X (N/2+1, nfram, nch)
complex64 ; istft_mc(X, hop, flag_noDiv=1, window=sqrt-hann) -> (xr (nch, nsampl'), N).  Only what the DR-NMF path
uses is implemented: the sqrt-Hann window of audio_dataset.py:194 and flag_noDiv=1 (audio_dataset.py:276)."""
from __future__ import annotations

import numpy as np
import torch

from . import engine as _engine


def sqrt_hann(N):
    """audio_dataset.py:194."""
    n = np.arange(N, dtype=np.float64)
    return np.sqrt((0.5 - 0.5 * np.cos(2.0 * np.pi * n / N)).astype(np.float32))


def _check_window(window, N):
    if window is not None and not np.allclose(np.asarray(window, np.float32), sqrt_hann(N), atol=1e-6):
        raise NotImplementedError("only the sqrt-Hann window of audio_dataset.py:194 is built into the kernels")


def stft_mc(x, N=1024, hop=None, window=None):
    """util.py:171-201."""
    if hop is None:
        hop = N // 2
    _check_window(window, N)
    x = np.asarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x.reshape(1, -1)
    nch, nsampl = x.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    audio = torch.as_tensor(np.ascontiguousarray(x).reshape(-1), device=dev)
    stack, _, fidx = _engine.stft_mag(audio, [c * nsampl for c in range(nch)], [nsampl] * nch, N, hop, want_mag=False)
    F = N // 2 + 1
    T = int(fidx[0, 1] - fidx[0, 0])
    s = stack.cpu().numpy()
    X = (s[:F] + 1j * s[F:]).astype(np.complex64).reshape(F, nch, T).transpose(0, 2, 1)
    return np.ascontiguousarray(X)


def istft_mc(X, hop, dtype=np.float32, nsampl=None, flag_noDiv=0, window=None):
    """util.py:203-226 (flag_noDiv=1 only: the librosa istft branch is never taken by the reference path)."""
    if not flag_noDiv:
        raise NotImplementedError("flag_noDiv=0 (librosa.istft with window-sum division) is not on the path")
    N = 2 * (X.shape[0] - 1)
    _check_window(window, N)
    F, T, nch = X.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    Xc = np.ascontiguousarray(X.transpose(0, 2, 1)).reshape(F, nch * T)
    stack = torch.as_tensor(np.concatenate([Xc.real, Xc.imag], axis=0).astype(np.float32), device=dev)
    fidx = torch.as_tensor(np.stack([np.arange(nch) * T, (np.arange(nch) + 1) * T], axis=1).astype(np.int64), device=dev)
    ys = _engine.mask_istft(stack, None, fidx, N, hop)
    xr = np.stack([y.cpu().numpy() for y in ys], axis=0).astype(dtype)
    if nsampl is not None:
        xr = xr[:, :nsampl]
    return xr, N


def masked_seqs_to_frames(x, mask):
    """util.py:19-27 (host-side reshaping)."""
    n_examples, time_steps, n_feature = x.shape
    xr = np.reshape(x.transpose((2, 0, 1)), (n_feature, n_examples * time_steps))
    m = np.reshape(mask.transpose((2, 0, 1)), (n_examples * time_steps,))
    return xr[:, np.where(m == m[0])[0]]
