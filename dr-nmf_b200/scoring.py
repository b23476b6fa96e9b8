"""Host-side scoring helpers that sit right after the hot path (score_audio.m:199-206): single-source BSS-Eval SDR.
Small numpy code on the host; the reference runs this step in a MATLAB subprocess with un-vendored toolboxes."""
import numpy as np


def sdr_db(est, ref):
    """s_t = (<est,ref>/|ref|^2) ref ; SDR = 10 log10(|s_t|^2 / |est - s_t|^2), both cut to the shorter length."""
    est = np.asarray(est, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    n = min(est.size, ref.size)
    est, ref = est[:n], ref[:n]
    st = (np.dot(est, ref) / np.dot(ref, ref)) * ref
    return 10.0 * np.log10(np.dot(st, st) / max(np.dot(est - st, est - st), 1e-300))


def wav_quantize(x):
    """util.py:37-45 / :29-35: float -> int16 wav -> float round trip."""
    x = np.asarray(x, dtype=np.float32)
    mx = np.max(np.abs(x)) if x.size else 0.0
    if mx > 1:
        x = x / mx
    return np.int16(x * 32767.0).astype(np.float32) / 32768.0
