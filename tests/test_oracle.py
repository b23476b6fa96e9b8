"""CPU tests: the numpy oracle against the committed golden vectors (tests/golden/*.npz were produced by
oracle/pin_reference.py, i.e. by the reference's own Python executed under numpy stand-ins) and against the
properties the reference's algorithms guarantee."""
import os

import numpy as np
import pytest

import oracle as O
from drnmf_b200 import synth


def _params(g, tag):
    return {k: g["%s_%s" % (tag, k)] for k in
            ("log_D", "log_alph", "log_lam1", "log_U1", "log_Uk", "log_h0", "k_clean", "k_noise")}


@pytest.mark.parametrize("tag", ["a", "b"])
def test_forward_matches_reference_run(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "drnmf_forward.npz"))
    p = _params(g, tag)
    x = g[tag + "_x"]
    H, irm = O.drnmf_forward(x, p, dtype=np.float64, dense_U=True)
    np.testing.assert_allclose(H, g[tag + "_H"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(irm, g[tag + "_irm"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(H[:, 0, :], g[tag + "_step0"], rtol=1e-11, atol=1e-13)
    # structured U (what the kernels compute) == dense U (what the reference multiplies by)
    Hs = O.rnn_forward(x, p, dtype=np.float64)
    np.testing.assert_allclose(Hs, g[tag + "_H"], rtol=1e-8, atol=1e-11)
    # the reference's float32 graph stays within the parity tolerance of the float64 truth
    H32 = O.rnn_forward(x, p, dtype=np.float32)
    assert np.linalg.norm(H32 - g[tag + "_H"]) / np.linalg.norm(g[tag + "_H"]) < 1e-5


def test_masked_frames_carry_state(golden_dir):
    g = np.load(os.path.join(golden_dir, "drnmf_forward.npz"))
    x, H = g["a_x"], g["a_H"]
    m = np.any(x != -1.0, axis=-1)
    for b in range(x.shape[0]):
        last = int(m[b].sum())
        assert last < x.shape[1] or b == 0
        for t in range(last, x.shape[1]):
            np.testing.assert_array_equal(H[b, t], H[b, last - 1])
    irm = g["a_irm"]
    assert np.all(irm > 0) and np.all(irm < 1)


def test_leading_masked_frame_outputs_zero_and_keeps_h0():
    rng = np.random.default_rng(3)
    F, R, K = 9, 6, 3
    p = O.alt_params_init(synth.dictionary(F, R), 5.0, 0.1, K, rng=rng)
    x = np.abs(rng.standard_normal((1, 4, F))).astype(np.float32)
    x[0, 0] = -1.0
    H = O.rnn_forward(x, p)
    assert np.all(H[0, 0] == 0.0)
    H_ref = O.rnn_forward(x[:, 1:], p)          # dropping the masked frame must give the same continuation
    np.testing.assert_allclose(H[:, 1:], H_ref, rtol=1e-13)


def test_ista_ed_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ista_ed.npz"))
    H = O.ista_ed(g["x"], g["W"], g["H0"].copy(), float(g["lam1"]), float(g["alph"]), int(g["K"]))
    np.testing.assert_allclose(H, g["H"], rtol=1e-13)


def test_unfolded_layers_equal_ista_when_tied():
    """With tied weights, h0-independent start and no leak, K layers of the network = K ISTA iterations from the
    previous frame's solution (enhance.py:144) -- ties oracle A.1 to oracle A.2."""
    rng = np.random.default_rng(11)
    F, R, K = 17, 8, 4
    W = synth.dictionary(F, R).astype(np.float64)
    alph, lam1 = 6.0, 0.2
    p = O.alt_params_init(W, alph, lam1, K, rng=rng)
    p["log_U1"] = np.log(np.eye(R) + 0.0).clip(-800)         # exact identity, no 1e-7 leak
    p["log_U1"][~np.eye(R, dtype=bool)] = -800.0
    p["log_Uk"] = np.full((R, R), -800.0)
    x = np.abs(rng.standard_normal((1, 1, F)))
    H = O.rnn_forward(x, p, dtype=np.float64, dense_U=True)[0, 0]
    Dn = np.exp(p["log_D"][0].astype(np.float64))
    Dn /= np.sqrt((Dn ** 2).sum(0, keepdims=True))
    a = float(np.exp(np.float64(p["log_alph"][0])))
    lam = float(np.exp(np.float64(p["log_lam1"][0])))
    h0 = np.logaddexp(p["log_h0"].astype(np.float64), 0)[:, None]
    # layer 0 has no Gram term: h <- relu(h0 + W^T x/alph - lam/alph); then K-1 full ISTA iterations
    h = np.maximum(0, h0 + Dn.T @ x[0].T / a - lam / a)
    h = O.ista_ed(x[0].T, Dn, h, lam, a, K - 1)
    np.testing.assert_allclose(H, h[:, 0], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_stft_istft_golden(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "stft_istft.npz"))
    N, hop, x = int(g[tag + "_N"]), int(g[tag + "_hop"]), g[tag + "_x"]
    win = O.sqrt_hann(N)
    X = O.stft_mc(x, N, hop, win)
    assert X.shape[1] == synth.stft_frames(len(x), N, hop)
    np.testing.assert_allclose(O.stack_reim(X), g[tag + "_stack"], atol=2e-5)
    np.testing.assert_allclose(O.reconstruct_x(g[tag + "_stack"], hop, win), g[tag + "_xr"], atol=1e-6)
    np.testing.assert_allclose(O.reconstruct_x(g[tag + "_stack"], hop, win, mask=g[tag + "_mask"]),
                               g[tag + "_xr_masked"], atol=1e-6)
    xr = O.reconstruct_x(g[tag + "_stack"], hop, win)[0, :len(x)]
    assert np.mean((x - xr) ** 2) / np.mean(x ** 2) < 1e-10     # test_audio_dataset.py:78-89's NMSE


def test_snmf_golden_and_properties(golden_dir):
    g = np.load(os.path.join(golden_dir, "snmf_ed.npz"))
    prm = {"cf": "ed", "sparsity": float(g["sparsity"]), "max_iter": int(g["max_iter"]),
           "conv_eps": float(g["conv_eps"]), "r": g["init_w"].shape[1], "init_w": g["init_w"].copy(),
           "w_update_ind": g["w_update_ind"], "init_h": "ones"}
    W, H, obj = O.sparse_nmf_chunked(g["V"], prm)
    np.testing.assert_allclose(W, g["W"], rtol=1e-12)
    np.testing.assert_allclose(H, g["H"], rtol=1e-12)
    np.testing.assert_allclose(obj["cost"], g["cost"], rtol=1e-12)
    # properties of the MU-ED algorithm (sparse_nmf_gpu.m): nonneg, unit-l2 columns, frozen atoms keep direction
    assert np.all(W >= 0) and np.all(H >= 0)
    np.testing.assert_allclose(np.sqrt((W ** 2).sum(0)), 1.0, rtol=1e-12)
    w0 = g["init_w"] / np.sqrt((g["init_w"] ** 2).sum(0))
    np.testing.assert_allclose(W[:, ~g["w_update_ind"]], w0[:, ~g["w_update_ind"]], rtol=1e-10)
    # all-updated run: cost is non-increasing
    prm2 = dict(prm); prm2.pop("w_update_ind"); prm2["conv_eps"] = 0.0; prm2["max_iter"] = 40
    _, _, o2 = O.sparse_nmf_ed(g["V"], prm2)
    assert np.all(np.diff(o2["cost"]) <= 1e-9 * o2["cost"][:-1])


def test_snmf_beta_divergence_branches_are_consistent():
    """sparse_nmf_gpu.m has four code paths (beta = 1, 2, 0-divergence, generic).  The MATLAB solver cannot run here
    [unpinned]; what can be checked is that the restated branches agree with one another where the formulas meet:
    the generic branch at beta -> 1 equals the KL branch, at beta -> 2 it gives the ED iterates with HALF the ED
    divergence (:271 has no 1/2, :275-276 divides by beta (beta - 1) = 2), at beta -> 0 the IS divergence; zeros of V
    are raised to its smallest positive entry for beta != 2 only (:201-205); KL cost is non-increasing."""
    rng = np.random.default_rng(11)
    F, n, R = 24, 90, 7
    V = np.abs(rng.standard_normal((F, n)))
    V[rng.random((F, n)) < 0.05] = 0.0
    base = {"sparsity": 0.3, "max_iter": 12, "conv_eps": 0.0, "r": R, "init_w": np.abs(rng.standard_normal((F, R))) + 0.1,
            "init_h": np.abs(rng.standard_normal((R, n))) + 0.1}
    eps = 1e-7
    Wk, Hk, ok = O.sparse_nmf_beta(V, dict(base, cf="kl"))
    Wg, Hg, og = O.sparse_nmf_beta(V, dict(base, cf="beta", beta=1.0 + eps))
    np.testing.assert_allclose(Wg, Wk, rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(Hg, Hk, rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(og["div"], ok["div"], rtol=1e-4)
    assert np.all(np.diff(ok["cost"]) <= 1e-9 * ok["cost"][:-1])
    Vp = V.copy(); Vp[Vp == 0] = Vp[Vp > 0].min()
    We, He, oe = O.sparse_nmf_beta(Vp, dict(base, cf="ed"))
    Wg, Hg, og = O.sparse_nmf_beta(Vp, dict(base, cf="beta", beta=2.0 + eps))
    np.testing.assert_allclose(Wg, We, rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(og["div"], 0.5 * oe["div"], rtol=1e-4)
    _, _, oe0 = O.sparse_nmf_beta(V, dict(base, cf="ed"))                 # ED keeps the zeros of V
    assert abs(oe0["div"][0] - oe["div"][0]) > 0
    Wi, Hi, oi = O.sparse_nmf_beta(V, dict(base, cf="is"))
    Wg, Hg, og = O.sparse_nmf_beta(V, dict(base, cf="beta", beta=eps))
    np.testing.assert_allclose(Wg, Wi, rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(og["div"], oi["div"], rtol=1e-4)
    assert O.sparse_nmf_beta(V, dict(base))[2]["div"][0] == ok["div"][0]   # cf defaults to 'kl' (:100-102)


def test_snmf_chunking_carries_dictionary():
    rng = np.random.default_rng(5)
    V = np.abs(rng.standard_normal((10, 50)))
    prm = {"sparsity": 0.1, "max_iter": 5, "r": 3, "init_w": np.abs(rng.standard_normal((10, 3))) + 0.1,
           "init_h": "ones"}
    W, H, obj = O.sparse_nmf_chunked(V, prm, frame_batch_size=20)
    assert len(obj["obj_snmf_per_chunk"]) == 3 and H.shape == (3, 50)
    # chunk 2 must start from chunk 1's dictionary (snmf.py:60-64)
    W1, _, _ = O.sparse_nmf_ed(V[:, :20], prm)
    p2 = dict(prm); p2["init_w"] = W1
    W2, _, _ = O.sparse_nmf_ed(V[:, 20:40], p2)
    p3 = dict(prm); p3["init_w"] = W2
    W3, _, _ = O.sparse_nmf_ed(V[:, 40:], p3)
    np.testing.assert_allclose(W, W3, rtol=1e-12)


def test_train_snmf_freezes_clean_atoms():
    rng = np.random.default_rng(8)
    F, n, r = 12, 40, 3
    clean, noisy = np.abs(rng.standard_normal((F, n))), np.abs(rng.standard_normal((F, n)))
    prm = {"cf": "ed", "sparsity": 0.2, "max_iter": 8, "conv_eps": 0.0, "r": r}
    Wc, _, _ = O.sparse_nmf_chunked(clean, dict(prm, init_w=np.abs(rng.standard_normal((F, r))) + 0.1, init_h="ones"))
    rng = np.random.default_rng(8)
    clean2, noisy2 = np.abs(rng.standard_normal((F, n))), np.abs(rng.standard_normal((F, n)))
    Wn, Hn, obj = O.train_snmf(clean2, noisy2, prm, noise_init=np.abs(rng.standard_normal((F, r))),
                               init_w_clean=np.abs(np.random.default_rng(8).standard_normal((F, r))) + 0.1)
    assert Wn.shape == (F, 2 * r) and Hn.shape == (2 * r, n)
    np.testing.assert_allclose(np.sqrt((Wn ** 2).sum(0)), 1.0, rtol=1e-12)


def test_param_count_identities():
    # plot_learning_curves_waspaa2017.ipynb:121-126
    assert O.param_count_notebook(257, 100, 2) == 103002
    assert O.param_count_notebook(257, 1000, 2) == 1030002
    assert O.param_count_notebook(257, 100, 5) == 257205
    assert O.param_count_notebook(257, 1000, 5) == 2572005


def test_sdr_and_wav_quantise():
    rng = np.random.default_rng(1)
    s = rng.standard_normal(4000)
    assert O.sdr_db(2.0 * s, s) > 200           # scale invariant
    e = s + 0.1 * rng.standard_normal(4000)
    assert 18 < O.sdr_db(e, s) < 22
    q = O.wav_quantize(np.float32(0.5) * np.ones(4))
    np.testing.assert_allclose(q, 16383 / 32768.0)


def test_synth_dictionary_is_contractive():
    F, R = 65, 40
    W = synth.dictionary(F, R).astype(np.float64)
    Dn = W / np.sqrt((W ** 2).sum(0, keepdims=True))
    lam_max = np.linalg.eigvalsh(Dn.T @ Dn)[-1]
    assert lam_max < R            # far below the random-dictionary 0.75 R; alph is chosen above it in tests
    n, c = synth.utterance(0, seconds=0.25)
    assert n.dtype == np.float32 and n.shape == (4000,) and abs(np.max(np.abs(c)) - 0.3) < 1e-6
