"""Host-side scoring helpers that sit right after the hot path (score_audio.m:177-238): single-source BSS-Eval SDR and
raw SNR.  Small numpy code on the host; the reference runs this step in a MATLAB subprocess with un-vendored toolboxes
(Voicebox snrseg, PESQ, STOI are not restated: their columns are reported as NaN)."""
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


def snr_db(est, ref):
    """score_audio.m:209: 10 log10(sum(ref^2) / sum((ref - est)^2)), both cut to the shorter length (:199-204)."""
    est = np.asarray(est, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    n = min(est.size, ref.size)
    est, ref = est[:n], ref[:n]
    return 10.0 * np.log10(np.dot(ref, ref) / max(np.dot(ref - est, ref - est), 1e-300))


SCORE_LABELS = ["SDR", "SNR", "SegSNR local", "SegSNR global", "PESQ", "STOI"]     # score_audio.m:233


def compute_scores(est, ref):
    """Row of score_audio.m:compute_scores for one (estimate, reference) pair of waveforms; the toolbox-only columns
    are NaN (the reference writes -1 for a skipped PESQ, :227)."""
    return np.array([sdr_db(est, ref), snr_db(est, ref), np.nan, np.nan, np.nan, np.nan]), list(SCORE_LABELS)


def print_scores(scores, labels, prefix=""):
    """enhance.py:355-359."""
    scores = np.atleast_2d(scores)
    for i, label in enumerate(labels):
        print("%sMean %s %.3f" % (prefix, label, np.nanmean(scores[:, i])))
