"""Synthetic CHiME2-shaped workload (there is no CHiME2 data offline).  numpy only; used by tests and bench.py for
BOTH arms so that the CUDA path and the CPU oracle see identical inputs.  Specification: SURVEY.md 8(d).

  audio       fs = 16 kHz; clean = 8..20 harmonics of an f0 gliding in 90..250 Hz, amplitude-modulated at 3..6 Hz,
              peak 0.3; noise = white Gaussian through a 1-pole low-pass (a = 0.95) scaled to 0 dB SNR;
              rng = default_rng(7654 + utterance index)        (7654 = the reference's seed, enhance.py:7)
  dictionary  per layer k: peaky spectral templates (1..5 Gaussian peaks, sigma 2.5 bins, + 1e-3) so that
              lambda_max(D^T D) <= alph and the unfolded ISTA is contractive; rng = default_rng(2016 + k)
              (2016 = the reference's SNMF seed, enhance.py:576)
"""
from __future__ import annotations

import math

import numpy as np

FS = 16000


def stft_frames(nsampl, N, hop):
    """util.py:184-190 + librosa center=False: ceil(n/hop) + N/hop + 1 frames."""
    nfram = int(math.ceil(float(nsampl) / float(hop)))
    return 1 + (nfram * hop + N) // hop


def utterance(index, seconds=3.0, fs=FS):
    """Returns (noisy, clean) float32 waveforms of one synthetic utterance."""
    rng = np.random.default_rng(7654 + int(index))
    n = int(round(seconds * fs))
    t = np.arange(n) / fs
    f0a, f0b = rng.uniform(90, 250, size=2)
    f0 = f0a + (f0b - f0a) * t / max(t[-1], 1e-9)
    phase = 2 * np.pi * np.cumsum(f0) / fs
    nh = int(rng.integers(8, 21))
    clean = np.zeros(n)
    for h in range(1, nh + 1):
        clean += rng.uniform(0.2, 1.0) / h * np.sin(h * phase + rng.uniform(0, 2 * np.pi))
    am = 0.55 + 0.45 * np.sin(2 * np.pi * rng.uniform(3, 6) * t + rng.uniform(0, 2 * np.pi))
    clean *= am
    clean *= 0.3 / np.max(np.abs(clean))
    w = rng.standard_normal(n)
    a = 0.95
    from scipy.signal import lfilter
    noise = lfilter([1.0 - a], [1.0, -a], w)                               # 1-pole low-pass
    noise *= np.sqrt(np.mean(clean ** 2) / np.mean(noise ** 2))          # 0 dB SNR
    noisy = clean + noise
    return noisy.astype(np.float32), clean.astype(np.float32)


def dictionary(F, R, k=0, seed=2016):
    """(F, R) nonnegative dictionary of peaky spectral templates; first R/2 atoms 'speech', last R/2 'noise'."""
    rng = np.random.default_rng(seed + int(k))
    f = np.arange(F, dtype=np.float64)[:, None]
    npk = rng.integers(1, 6, size=R)
    W = np.full((F, R), 1e-3)
    sigma = 2.5
    for j in range(R):
        c = rng.uniform(0, F, size=npk[j])
        hgt = rng.uniform(0.2, 1.2, size=npk[j])
        W[:, j] += (hgt[None, :] * np.exp(-0.5 * ((f - c[None, :]) / sigma) ** 2)).sum(axis=1)
    return W.astype(np.float32)


def default_alph(R):
    """enhance.py:608-614 defines 50/200/400 for r = 100/500/1000; extrapolated linearly in R beyond."""
    return {200: 50.0, 1000: 200.0, 2000: 400.0}.get(int(R), max(50.0, 0.2 * R))


def model_params(F, R, K, alph=None, lam1=1.0, seed=2016, untied=True):
    """Parameter dict in the layout of DrnmfEngine.set_params / the oracle (untied per-layer dictionaries)."""
    if alph is None:
        alph = default_alph(R)
    eps32 = np.float32(1e-7)
    r = R // 2
    Ws = [dictionary(F, R, k if untied else 0, seed) for k in range(K)]
    log_D = np.stack([np.log(1e-7 + w).astype(np.float32) for w in Ws])
    rng = np.random.default_rng(seed + 1000)
    W0 = Ws[0]
    return {
        "log_D": log_D,
        "log_alph": np.full((K,), np.log(eps32 + np.float32(alph)), dtype=np.float32),
        "log_lam1": np.full((K,), np.log(eps32 + np.float32(lam1)), dtype=np.float32),
        "log_U1": np.log(eps32 + np.eye(R, dtype=np.float32)),
        "log_Uk": np.log(eps32 + np.zeros((R, R), dtype=np.float32)),
        "log_h0": rng.uniform(-0.05, 0.05, size=(R,)).astype(np.float32),
        "k_clean": np.log(1e-7 + W0[:, :r]).astype(np.float32).T.copy(),
        "k_noise": np.log(1e-7 + W0[:, r:]).astype(np.float32).T.copy(),
    }


def structured_u_init():
    """(diag, off) pairs of exp(log_U1)^T / exp(log_Uk)^T at build_alt's initial values, without the R x R arrays."""
    eps32 = np.float32(1e-7)
    d0 = float(np.exp(np.log(eps32 + np.float32(1.0))))
    o = float(np.exp(np.log(eps32 + np.float32(0.0))))
    return (d0, o), (o, o)
