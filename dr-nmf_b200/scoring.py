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


def snrseg(est, ref, fs, tf=0.01):
    """Segmental and global SNR in dB as score_audio.m:211-212 asks of Voicebox: [loc, glo] = snrseg(xest, xref, fs).
    Restated from the published description of v_snrseg: non-overlapping frames of tf = 10 ms, per-frame
    10 log10(sum ref^2 / sum (ref - est)^2) clipped to [-10, 35] dB and averaged; global = the same ratio over all frames.
    [unpinned] Voicebox is not vendored in the reference.  Its default mode 'Vq' additionally drops frames that the
    ITU-T P.56 activity detector marks silent and removes +-1-sample delays by quadratic interpolation; this is mode
    'wz' (no VAD, no alignment): equal on signals that are active throughout and time-aligned."""
    est = np.asarray(est, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    n = min(est.size, ref.size)
    L = int(round(tf * fs))
    nf = n // L
    if nf == 0:
        return np.nan, np.nan
    e = (ref[:nf * L] - est[:nf * L]).reshape(nf, L)
    r = ref[:nf * L].reshape(nf, L)
    pe, pr = np.sum(e * e, axis=1), np.sum(r * r, axis=1)
    keep = pr > 0
    with np.errstate(divide="ignore"):
        seg = 10.0 * np.log10(pr[keep] / np.maximum(pe[keep], 1e-300))
    loc = float(np.mean(np.clip(seg, -10.0, 35.0))) if seg.size else np.nan
    glo = float(10.0 * np.log10(np.sum(pr) / max(np.sum(pe), 1e-300)))
    return loc, glo


def compute_scores(est, ref, fs=16000):
    """Row of score_audio.m:compute_scores for one (estimate, reference) pair of waveforms: SDR (BSS-Eval, single
    source), raw SNR, segmental SNR (see snrseg); PESQ and STOI are toolbox code that is not vendored: NaN (the reference
    writes -1 for a skipped PESQ, :227)."""
    loc, glo = snrseg(est, ref, fs)
    return np.array([sdr_db(est, ref), snr_db(est, ref), loc, glo, np.nan, np.nan]), list(SCORE_LABELS)


def aggregate_scores(scores_per_snr, labels=None, scores_to_print=("SDR",), print_per_snr=True):
    """The numeric part of print_scores.py:print_row (:84-114): `scores_per_snr` maps an SNR label (e.g. 'm6dB') to an
    (n_files, n_scores) array; returns (dict score -> {'per_snr': {snr: mean}, 'all': mean over every file}, LaTeX row
    fragment '%.2f & ... \\\\' in the reference's column order: per-SNR means first, then the mean over SNRs)."""
    labels = list(SCORE_LABELS) if labels is None else list(labels)
    snrs = list(scores_per_snr.keys())
    allrows = np.concatenate([np.atleast_2d(scores_per_snr[s]) for s in snrs], axis=0)
    out, row = {}, ""
    for i, lab in enumerate(labels):
        if lab not in scores_to_print:
            continue
        per = {s: float(np.mean(np.atleast_2d(scores_per_snr[s])[:, i])) for s in snrs}
        out[lab] = {"per_snr": per, "all": float(np.mean(allrows[:, i]))}
        if print_per_snr:
            for s in snrs:
                row += "%.2f & " % per[s]
        row += "%.2f & " % out[lab]["all"]
    return out, row[:-3] + " \\\\"


def print_scores(scores, labels, prefix=""):
    """enhance.py:355-359."""
    scores = np.atleast_2d(scores)
    for i, label in enumerate(labels):
        print("%sMean %s %.3f" % (prefix, label, np.nanmean(scores[:, i])))
