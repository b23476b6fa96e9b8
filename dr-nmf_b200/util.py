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


def wavread(wavfile):
    """util.py:29-35: int16 wav -> float32 (nch, nsampl) scaled by 1/32768 (a list argument means its first entry)."""
    import scipy.io.wavfile
    if isinstance(wavfile, list):
        wavfile = wavfile[0]
    fs, x = scipy.io.wavfile.read(wavfile)
    x = np.transpose(x).astype(np.float32) / np.float32(32768.0)
    return x.reshape(1, -1) if x.ndim == 1 else x


def wavwrite(wavfile, fs, x):
    """util.py:37-45: x (nch, nsampl); float32 data is peak-normalised only when it clips, then cast to int16 (x * 32767)."""
    import scipy.io.wavfile
    x = np.asarray(x)
    if x.dtype == np.float32:
        mx = np.max(np.abs(x)) if x.size else 0.0
        if mx > 1:
            x = x / mx
        x = np.int16(x * 32767.0)
    scipy.io.wavfile.write(wavfile, fs, x.T)


def compute_STFTs(wavfiles, params_stft):
    """util.py:327-352: [Re;Im] stack (2F, total frames) of all files (first `nch` channels... here channel 0, as
    params_stft['nch'] = 1 in every shipped config) and the (n_files, 2) frame index table, computed by the CUDA STFT."""
    waves = [wavread(f)[0] for f in wavfiles]
    lens = [len(w) for w in waves]
    offs = np.concatenate([[0], np.cumsum(lens)])[:-1]
    dev = torch.device("cuda", torch.cuda.current_device())
    audio = torch.as_tensor(np.concatenate(waves).astype(np.float32), device=dev)
    stack, _, fidx = _engine.stft_mag(audio, list(offs), lens, int(params_stft["N"]), int(params_stft["hop"]), want_mag=False)
    return stack.cpu().numpy(), fidx.cpu().numpy().astype(np.int32)
