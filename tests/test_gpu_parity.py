"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C-ABI, against the numpy oracle
(float64 truth) on the same seeded inputs and against the committed golden fixtures.

Tolerances are the ones BASELINE.json's north_star states: relative error <= 1e-4 on H and on the mask
(Frobenius AND max-abs/max), <= 0.01 dB on the SDR of the reconstructed audio.
"""
import os
import sys

import numpy as np
import pytest
import torch

import oracle as O
from drnmf_b200 import engine, synth

pytestmark = pytest.mark.gpu

TOL = 1e-4           # north_star: relative error <= 1e-4 on H and on the mask
IMPLS = ["simt", "tc"]


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    fro = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
    mx = np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)
    return fro, mx


def _golden_params(g, tag):
    return {k: g["%s_%s" % (tag, k)] for k in
            ("log_D", "log_alph", "log_lam1", "log_U1", "log_Uk", "log_h0", "k_clean", "k_noise")}


def _run(p, x, impl, square=False):
    K, F, R = p["log_D"].shape
    eng = engine.DrnmfEngine(F, R, K, square_irm=square, impl=impl)
    eng.set_params(p)
    H, irm = eng.forward(torch.as_tensor(x, device="cuda"))
    torch.cuda.synchronize()
    if impl == "tc" and os.environ.get("DRNMF_RECURRENT") != "simt":
        # the product path must really be the persistent tcgen05 kernel, not the CUDA-core lock path
        assert eng.recurrent_config()["impl"] == "tcgen05", eng.recurrent_config()
    return eng, H.cpu().numpy(), irm.cpu().numpy()


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("tag", ["a", "b"])
def test_forward_golden(golden_dir, tag, impl):
    """Fixtures produced by the reference's own step()/build_alt code (oracle/pin_reference.py)."""
    g = np.load(os.path.join(golden_dir, "drnmf_forward.npz"))
    p = _golden_params(g, tag)
    _, H, irm = _run(p, g[tag + "_x"], impl)
    for got, want in ((H, g[tag + "_H"]), (irm, g[tag + "_irm"])):
        fro, mx = rel_err(got, want)
        assert fro < TOL and mx < TOL, (impl, tag, fro, mx)


@pytest.mark.parametrize("impl", IMPLS)
def test_derived_weights(impl):
    F, R, K = 70, 48, 3
    p = synth.model_params(F, R, K, alph=20.0)
    p["log_alph"] = (p["log_alph"] + np.array([0.0, 0.1, -0.2], np.float32)).astype(np.float32)
    eng = engine.DrnmfEngine(F, R, K, impl=impl)
    eng.set_params(p)
    for k in range(K):
        Wk, Sk, bk = O.layer_weights(p, k)
        WT = eng.derived(1, k).cpu().numpy()
        assert rel_err(WT[:R, :F], Wk.T)[1] < 1e-5
        assert np.all(WT[R:] == 0) and np.all(WT[:, F:] == 0)
        assert rel_err(eng.derived(2, k).cpu().numpy()[:R], bk)[1] < 1e-6
        if k > 0:
            ST = eng.derived(0, k).cpu().numpy()
            assert rel_err(ST[:R, :R], Sk.T)[1] < 2e-6, (impl, k, rel_err(ST[:R, :R], Sk.T))
            assert np.all(ST[R:] == 0) and np.all(ST[:, R:] == 0)
    h0 = np.logaddexp(p["log_h0"].astype(np.float64), 0)
    assert rel_err(eng.derived(3).cpu().numpy()[:R], h0)[1] < 1e-6


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("shape", [
    dict(F=65, R=40, K=4, B=5, T=12, alph=30.0),            # ragged lengths, R not a multiple of 8
    dict(F=129, R=200, K=5, B=3, T=9, alph=50.0),           # reference-like r=100 (Rp=256)
    dict(F=257, R=200, K=2, B=70, T=6, alph=50.0),          # B above one batch tile
    dict(F=33, R=16, K=1, B=2, T=5, alph=10.0),             # single layer: no Gram term at all
    dict(F=65, R=1200, K=3, B=5, T=4, alph=300.0),          # R > 1024: padded to 1536, weights stream through TMEM in chunks
    dict(F=40, R=600, K=3, B=3, T=3, alph=150.0),           # 5 x 128 atoms has no tiling: padded to 768
])
def test_forward_vs_oracle(shape, impl):
    F, R, K, B, T = (shape[k] for k in "FRKBT")
    rng = np.random.default_rng(1234 + F + R)
    p = synth.model_params(F, R, K, alph=shape["alph"], lam1=0.5)
    p["log_alph"] = (p["log_alph"] + 0.05 * rng.standard_normal(K)).astype(np.float32)
    x = (np.abs(rng.standard_normal((B, T, F))) * 3.0).astype(np.float32)
    lens = rng.integers(1, T + 1, size=B)
    lens[0] = T
    for b in range(B):
        x[b, lens[b]:] = -1.0
    Ho, irmo = O.drnmf_forward(x, p, dtype=np.float64)
    _, H, irm = _run(p, x, impl)
    fro, mx = rel_err(H, Ho)
    assert fro < TOL and mx < TOL, ("H", impl, shape, fro, mx)
    fro, mx = rel_err(irm, irmo)
    assert fro < TOL and mx < TOL, ("irm", impl, shape, fro, mx)
    # masked frames carry the state; mask in (0,1)
    for b in range(B):
        for t in range(lens[b], T):
            np.testing.assert_array_equal(H[b, t], H[b, lens[b] - 1])
    assert np.all(irm > 0) and np.all(irm <= 1)


def test_north_star_shape():
    """F=513, R=1000 (r=500), K=25 as in BASELINE.json, short utterances so that the float64 oracle stays fast."""
    F, R, K, B, T = 513, 1000, 25, 6, 12
    rng = np.random.default_rng(2017)
    p = synth.model_params(F, R, K)
    x = (np.abs(rng.standard_normal((B, T, F))) * 4.0).astype(np.float32)
    x[1, 9:] = -1.0
    Ho, irmo = O.drnmf_forward(x, p, dtype=np.float64)
    eng, H, irm = _run(p, x, "tc")
    cfg = eng.recurrent_config()
    assert cfg["impl"] == "tcgen05" and cfg["MT"] == 8, cfg
    fro, mx = rel_err(H, Ho)
    assert fro < TOL and mx < TOL, ("H", fro, mx, cfg)
    fro, mx = rel_err(irm, irmo)
    assert fro < TOL and mx < TOL, ("irm", fro, mx, cfg)


@pytest.mark.parametrize("impl", IMPLS)
def test_untie_alph_and_square_irm(impl):
    F, R, K, B, T = 40, 24, 3, 2, 7
    rng = np.random.default_rng(77)
    p = synth.model_params(F, R, K, alph=15.0)
    p["log_alph"] = (np.log(15.0) + 0.1 * rng.standard_normal((K, R))).astype(np.float32)     # enhance.py:225-226
    x = (np.abs(rng.standard_normal((B, T, F))) * 2.0).astype(np.float32)
    Ho, irmo = O.drnmf_forward(x, p, dtype=np.float64, transform_before_irm="square")
    _, H, irm = _run(p, x, impl, square=True)
    assert max(rel_err(H, Ho)) < TOL and max(rel_err(irm, irmo)) < TOL


@pytest.mark.parametrize("impl", IMPLS)
def test_leading_masked_frame(impl):
    F, R, K = 20, 16, 3
    rng = np.random.default_rng(5)
    p = synth.model_params(F, R, K, alph=10.0)
    x = np.abs(rng.standard_normal((2, 5, F))).astype(np.float32)
    x[0, 0] = -1.0
    Ho = O.rnn_forward(x, p)
    _, H, _ = _run(p, x, impl)
    assert np.all(H[0, 0] == 0.0)
    assert max(rel_err(H, Ho)) < TOL


def test_tc_matches_simt_bitwise_stable():
    """Two runs of the tensor-core path give identical bits (deterministic reductions, no atomics)."""
    F, R, K, B, T = 65, 40, 4, 6, 8
    rng = np.random.default_rng(9)
    p = synth.model_params(F, R, K, alph=30.0)
    x = np.abs(rng.standard_normal((B, T, F))).astype(np.float32)
    eng, H1, irm1 = _run(p, x, "tc")
    H2, irm2 = eng.forward(torch.as_tensor(x, device="cuda"))
    assert np.array_equal(H1, H2.cpu().numpy()) and np.array_equal(irm1, irm2.cpu().numpy())


# ---- STFT / iSTFT ------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_stft_istft_golden(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "stft_istft.npz"))
    N, hop, x = int(g[tag + "_N"]), int(g[tag + "_hop"]), g[tag + "_x"]
    audio = torch.as_tensor(x, device="cuda")
    stack, mag, fidx = engine.stft_mag(audio, [0], [len(x)], N, hop)
    ref = g[tag + "_stack"]
    assert stack.shape == ref.shape
    assert np.max(np.abs(stack.cpu().numpy() - ref)) < 2e-5 * max(1.0, np.abs(ref).max())
    F = N // 2 + 1
    np.testing.assert_allclose(mag.cpu().numpy().T, np.sqrt(ref[:F] ** 2 + ref[F:] ** 2), atol=3e-5)
    ref_stack = torch.as_tensor(ref, device="cuda")
    for mask, want in ((None, g[tag + "_xr"]), (g[tag + "_mask"], g[tag + "_xr_masked"])):
        m = None if mask is None else torch.as_tensor(np.ascontiguousarray(mask.T), device="cuda")
        (y,) = engine.mask_istft(ref_stack, m, fidx, N, hop)
        assert y.numel() == want.shape[1]
        np.testing.assert_allclose(y.cpu().numpy(), want[0], atol=2e-6)


def test_stft_roundtrip_full_size():
    """BASELINE-sized property: STFT -> iSTFT reconstructs the signal (test_audio_dataset.py:78-89), ragged batch."""
    N, hop = 1024, 256
    lens = [48000, 32000, 40123, 300, 700]
    sigs = [synth.utterance(i, seconds=n / 16000.0)[0][:n] for i, n in enumerate(lens)]
    offs = np.concatenate([[0], np.cumsum(lens)])[:-1]
    audio = torch.as_tensor(np.concatenate(sigs), device="cuda")
    stack, mag, fidx = engine.stft_mag(audio, list(offs), lens, N, hop)
    ys = engine.mask_istft(stack, None, fidx, N, hop)
    for s, y, n in zip(sigs, ys, lens):
        y = y.cpu().numpy()[:n]
        assert np.mean((s - y) ** 2) / max(np.mean(s ** 2), 1e-30) < 1e-10
    # linearity of the masked synthesis: mask 0.5 halves the output
    half = torch.full_like(mag, 0.5)
    yh = engine.mask_istft(stack, half, fidx, N, hop)
    np.testing.assert_allclose(yh[0].cpu().numpy(), 0.5 * ys[0].cpu().numpy(), atol=1e-6)


@pytest.mark.parametrize("N,hop", [(128, 32), (256, 128), (512, 128), (1024, 256), (1024, 512), (2048, 512), (4096, 1024)])
def test_stft_sizes_vs_oracle(N, hop):
    """Every FFT path (four-step register transform for N = 512 / 1024, tiled radix-2, one-frame-per-CTA fallback for
    N = 4096) against the numpy oracle: stack, magnitude and masked synthesis, ragged batch incl. an odd frame count."""
    rng = np.random.default_rng(N + hop)
    lens = [int(3.3 * N) + 17, 9 * N + 5, 6 * N]
    sigs = [rng.standard_normal(n).astype(np.float32) for n in lens]
    offs = np.concatenate([[0], np.cumsum(lens)])[:-1]
    audio = torch.as_tensor(np.concatenate(sigs), device="cuda")
    stack, mag, fidx = engine.stft_mag(audio, list(offs), lens, N, hop)
    win = O.sqrt_hann(N)
    F = N // 2 + 1
    col = 0
    masks = []
    for s_ in sigs:
        ref = O.stack_reim(O.stft_mc(s_.reshape(1, -1), N, hop, win))
        T = ref.shape[1]
        got = stack[:, col:col + T].cpu().numpy()
        assert np.max(np.abs(got - ref)) < 3e-5 * max(1.0, np.abs(ref).max())
        np.testing.assert_allclose(mag[col:col + T].cpu().numpy().T, np.sqrt(ref[:F] ** 2 + ref[F:] ** 2), atol=1e-4, rtol=1e-5)
        masks.append(rng.random((T, F)).astype(np.float32))
        col += T
    assert col == stack.shape[1]
    m = torch.as_tensor(np.concatenate(masks), device="cuda")
    ys = engine.mask_istft(stack, m, fidx, N, hop)
    for s_, mk, y in zip(sigs, masks, ys):
        ref_stack = O.stack_reim(O.stft_mc(s_.reshape(1, -1), N, hop, win)).astype(np.float64)
        want = O.reconstruct_x(ref_stack, hop, win.astype(np.float64), mask=mk.T.astype(np.float64), dtype=np.float64)[0]
        got = y.cpu().numpy()
        assert got.size == want.size and got.size > 0
        np.testing.assert_allclose(got, want, atol=5e-6 * max(1.0, np.abs(want).max()))


# ---- end to end: SDR parity ---------------------------------------------------------------------
@pytest.mark.parametrize("impl", IMPLS)
def test_enhance_sdr_parity(impl):
    """north_star: <= 0.01 dB on the SDR of the reconstructed audio (float audio and after wav int16 quantisation)."""
    N, hop, R, K, B = 256, 64, 64, 4, 3
    F = N // 2 + 1
    secs = [0.5, 0.4, 0.3]
    pairs = [synth.utterance(i, seconds=s) for i, s in enumerate(secs)]
    win = O.sqrt_hann(N)
    frames = [synth.stft_frames(len(n), N, hop) for n, _ in pairs]
    T = max(frames)
    p = synth.model_params(F, R, K, alph=25.0)
    x = np.full((B, T, F), -1.0, np.float32)
    stack = np.zeros((2 * F, B * T), np.float32)
    for b, (noisy, _) in enumerate(pairs):
        Y = O.stack_reim(O.stft_mc(noisy, N, hop, win))
        x[b, :frames[b]] = O.magnitude(Y).T
        stack[:, b * T:b * T + frames[b]] = Y
    _, irm = O.drnmf_forward(x, p, dtype=np.float64)
    eng = engine.DrnmfEngine(F, R, K, impl=impl)
    eng.set_params(p)
    out = eng.enhance_host(torch.as_tensor(x).pin_memory(), torch.as_tensor(stack).pin_memory(),
                           torch.as_tensor(np.asarray(frames, np.int32)), N, hop).numpy()
    for b, (noisy, clean) in enumerate(pairs):
        ref = O.reconstruct_x(stack[:, b * T:b * T + frames[b]].astype(np.float64), hop, win.astype(np.float64),
                              mask=irm[b, :frames[b]].T, dtype=np.float64)[0]
        got = out[b, :ref.size]
        n = len(clean)
        d = abs(O.sdr_db(got[:n], clean) - O.sdr_db(ref[:n], clean))
        dq = abs(O.sdr_db(O.wav_quantize(got[:n]), clean) - O.sdr_db(O.wav_quantize(ref[:n].astype(np.float32)), clean))
        assert d < 0.01 and dq < 0.01, (impl, b, d, dq)
        assert np.all(out[b, ref.size:] == 0)


# ---- the reference-facing interfaces (custom_layers / enhance / util / audio_dataset mirrors) ---------------
def _build_params(F, r, K, T, W, **extra):
    p = {"input_dim": F, "hidden_dim": 2 * r, "output_dim": F, "mask_value": -1.0, "maxseq": T, "K_layers": K, "W": W,
         "alph": 20.0, "lam1": 0.5, "params_untied": ["log_D", "log_alph"], "params_trainable": ["log_D", "log_alph"]}
    p.update(extra)
    return p


@pytest.mark.parametrize("untie_alph", [False, True])
def test_build_unfolded_snmf_predict_on_batch(untie_alph):
    from drnmf_b200 import enhance
    F, r, K, B, T = 65, 12, 3, 4, 9
    rng = np.random.default_rng(31)
    W = synth.dictionary(F, 2 * r)
    model = enhance.build_unfolded_snmf(_build_params(F, r, K, T, W, untie_alph=untie_alph))
    x = (np.abs(rng.standard_normal((B, T, F))) * 2).astype(np.float32)
    x[2, 5:] = -1.0
    irm, H = model.predict_on_batch(x, return_hidden=True)
    # oracle with the SAME initial parameters (log_h0 is drawn by the layer -> read it back)
    p = O.alt_params_init(W, 20.0, 0.5, K, untie_alph=untie_alph)
    p["log_h0"] = model.rnn.log_h0.cpu().numpy()
    Ho, irmo = O.drnmf_forward(x, p, dtype=np.float64)
    assert max(rel_err(H, Ho)) < TOL and max(rel_err(irm, irmo)) < TOL
    # Keras-like weight surface: names, round trip through save/load
    names = model.weight_names()
    assert names[0].endswith("_log_h0") and any(n.endswith("_log_D_2") for n in names) and len(names) == len(model.get_weights())
    w = model.get_weights()
    model2 = enhance.build_unfolded_snmf(_build_params(F, r, K, T, np.ones_like(W), untie_alph=untie_alph))
    model2.set_weights(w)
    np.testing.assert_array_equal(model2.predict_on_batch(x), irm)


def test_util_and_audio_dataset_mirrors(golden_dir):
    from drnmf_b200 import audio_dataset, util
    g = np.load(os.path.join(golden_dir, "stft_istft.npz"))
    N, hop, x = int(g["b_N"]), int(g["b_hop"]), g["b_x"]
    win = util.sqrt_hann(N)
    X = util.stft_mc(x.reshape(1, -1), N, hop, win)
    F = N // 2 + 1
    np.testing.assert_allclose(np.concatenate([X[:, :, 0].real, X[:, :, 0].imag]), g["b_stack"], atol=3e-5)
    xr, n_out = util.istft_mc(X, hop, flag_noDiv=1, window=win)
    assert n_out == N
    np.testing.assert_allclose(xr, g["b_xr"], atol=3e-6)
    ds = audio_dataset.AudioDataset([x], params_stft={"N": N, "hop": hop, "nch": 1})
    np.testing.assert_allclose(ds.reconstruct_x(0, mask=g["b_mask"]), g["b_xr_masked"], atol=3e-6)
    with pytest.raises(NotImplementedError):
        util.stft_mc(x, N, hop, np.ones(N, np.float32))


def test_enhance_main_synthetic(tmp_path, capsys):
    from drnmf_b200 import enhance
    cfg = tmp_path / "params_unfolded_snmf_test.yaml"
    cfg.write_text("K_layers: 2\nalph: 50.0\nlam1: 1.0\nr: 20\nparams_trainable: [log_D, log_alph]\n"
                   "params_untied: [log_D, log_alph]\n")
    dat = tmp_path / "params_data.yaml"
    dat.write_text("maxlen: 500\nparams_stft: {N: 256, hop: 64, nch: 1}\n")
    assert enhance.main(["-c", str(cfg), "-d", str(dat), "--synthetic", "3", "--seconds", "0.4"]) == 0
    assert "mean SDR" in capsys.readouterr().out


# ---- sparse NMF multiplicative updates (sparse_nmf_gpu.m, ED branch) -----------------------------------------
@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("case", [dict(F=33, n=200, R=12, it=25, partial=False), dict(F=65, n=333, R=40, it=15, partial=True)])
def test_snmf_mu_ed_vs_oracle(case, impl):
    F, n, R, iters = case["F"], case["n"], case["R"], case["it"]
    rng = np.random.default_rng(2016 + F)
    V = (np.abs(rng.standard_normal((F, n))) * 2).astype(np.float32)
    W0 = (np.abs(rng.standard_normal((F, R))) + 0.1).astype(np.float32)
    H0 = (np.abs(rng.standard_normal((R, n))) + 0.1).astype(np.float32)
    prm = {"cf": "ed", "sparsity": 0.7, "max_iter": iters, "conv_eps": 0.0, "r": R, "init_w": W0, "init_h": H0}
    wu = None
    if case["partial"]:
        wu = np.arange(R) >= R // 2                 # first half frozen (train_snmf stage 2, enhance.py:110-116)
        prm["w_update_ind"] = wu
    Wo, Ho, obj = O.sparse_nmf_ed(V, prm, dtype=np.float64)
    Vd, Wd, Hd = (torch.as_tensor(a, device="cuda") for a in (V, W0.copy(), H0.copy()))
    cost, div = engine.snmf_mu_ed(Vd, Wd, Hd, 0.7, iters, 0.0, w_update=wu, impl=None if impl == "tc" else "simt")
    assert len(cost) == iters
    assert max(rel_err(Wd.cpu().numpy(), Wo)) < TOL, rel_err(Wd.cpu().numpy(), Wo)
    assert max(rel_err(Hd.cpu().numpy(), Ho)) < TOL, rel_err(Hd.cpu().numpy(), Ho)
    np.testing.assert_allclose(cost, obj["cost"], rtol=2e-5)
    np.testing.assert_allclose(div, obj["div"], rtol=2e-5)
    # properties of the algorithm: nonneg, unit-l2 columns, monotone cost when everything is updated
    Wn = Wd.cpu().numpy()
    assert np.all(Wn >= 0) and np.all(Hd.cpu().numpy() >= 0)
    np.testing.assert_allclose(np.sqrt((Wn.astype(np.float64) ** 2).sum(0)), 1.0, rtol=1e-5)
    if not case["partial"]:
        assert np.all(np.diff(cost) <= 1e-6 * cost[:-1])


def test_snmf_inference_mode_and_convergence():
    """W frozen (enhance.py:838-845): only H moves; conv_eps stops early exactly like sparse_nmf_gpu.m:288-296."""
    F, n, R = 40, 150, 16
    rng = np.random.default_rng(5)
    V = np.abs(rng.standard_normal((F, n))).astype(np.float32)
    W0 = (np.abs(rng.standard_normal((F, R))) + 0.1).astype(np.float32)
    prm = {"cf": "ed", "sparsity": 0.2, "max_iter": 200, "conv_eps": 1e-3, "r": R, "init_w": W0, "init_h": "ones",
           "w_update_ind": np.zeros(R, bool)}
    Wo, Ho, obj = O.sparse_nmf_ed(V, prm, dtype=np.float64)
    Vd, Wd = torch.as_tensor(V, device="cuda"), torch.as_tensor(W0.copy(), device="cuda")
    Hd = torch.ones((R, n), device="cuda")
    cost, div = engine.snmf_mu_ed(Vd, Wd, Hd, 0.2, 200, 1e-3, w_update=np.zeros(R, bool))
    assert len(cost) == len(obj["cost"]) < 200
    assert max(rel_err(Hd.cpu().numpy(), Ho)) < TOL
    W0n = W0 / np.sqrt((W0.astype(np.float64) ** 2).sum(0))
    assert max(rel_err(Wd.cpu().numpy(), W0n)) < 1e-6


def test_snmf_python_mirror_chunk_driver(golden_dir):
    """snmf.sparse_nmf_matlab (chunk driver + parameter handling) on the golden fixture of the reference's driver."""
    from drnmf_b200 import snmf
    g = np.load(os.path.join(golden_dir, "snmf_ed.npz"))
    prm = {"cf": "ed", "sparsity": float(g["sparsity"]), "max_iter": float(g["max_iter"]), "conv_eps": float(g["conv_eps"]),
           "display": 0., "random_seed": 2016., "r": g["init_w"].shape[1], "init_w": g["init_w"].copy(),
           "w_update_ind": g["w_update_ind"], "init_h": "ones"}
    W, H, obj = snmf.sparse_nmf_matlab(g["V"], prm, verbose=False)
    assert max(rel_err(W, g["W"])) < TOL and max(rel_err(H, g["H"])) < TOL
    np.testing.assert_allclose(obj["cost"], g["cost"], rtol=2e-5)
    with pytest.raises(NotImplementedError):                       # there is no CPU solver to fall back to
        snmf.sparse_nmf_matlab_on_chunk(g["V"], prm, useGPU=False)


# ---- training: loss + hand-written BPTT against torch.autograd on the float64 oracle ---------------------------
@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("case", [dict(F=33, R=16, K=3, B=3, T=6, tied=False), dict(F=65, R=40, K=4, B=5, T=8, tied=False),
                                  dict(F=40, R=24, K=3, B=2, T=5, tied=True),
                                  dict(F=33, R=24, K=4, B=3, T=6, tied=False, vec=True)])     # untie_alph (enhance.py:225-226)
def test_loss_and_grads_vs_autograd(case, impl):
    from oracle import torch_oracle as TO
    F, R, K, B, T = (case[k] for k in "FRKBT")
    rng = np.random.default_rng(99 + F)
    p = synth.model_params(F, R, K, alph=15.0, lam1=0.3, untied=not case["tied"])
    if not case["tied"]:
        p["log_alph"] = (p["log_alph"] + 0.05 * rng.standard_normal(K)).astype(np.float32)
    if case.get("vec"):   # one step size per atom: S_k is no longer symmetric, the backward chain must cope
        p["log_alph"] = (p["log_alph"][:, None] + 0.15 * rng.standard_normal((K, R))).astype(np.float32)
    x = (np.abs(rng.standard_normal((B, T, F))) * 2).astype(np.float32)
    y = (x * rng.uniform(0.2, 0.9, size=x.shape)).astype(np.float32)
    lens = rng.integers(2, T + 1, size=B); lens[0] = T
    for b in range(B):
        x[b, lens[b]:] = -1.0; y[b, lens[b]:] = -1.0
    loss_o, g_o, H_o, irm_o = TO.loss_and_grads(x, y, p)
    pe = dict(p)
    if case["tied"]:     # one shared dictionary / step size: gradients are summed over the layers
        pe["log_D"], pe["log_alph"], pe["log_lam1"] = p["log_D"][:1], p["log_alph"][:1], p["log_lam1"][:1]
        for k in ("log_D", "log_alph", "log_lam1"):
            g_o[k] = g_o[k].sum(axis=0, keepdims=True)
        assert np.all(p["log_D"] == p["log_D"][0])
    eng = engine.DrnmfEngine(F, R, K, impl=impl)
    eng.set_params(pe)
    ls, ms, g, irm = eng.loss_and_grads(torch.as_tensor(x, device="cuda"), torch.as_tensor(y, device="cuda"), want_irm=True)
    assert ms == float(lens.sum())
    assert abs(ls / ms - loss_o) < 2e-5 * abs(loss_o)
    assert max(rel_err(irm.cpu().numpy(), irm_o)) < TOL
    for key in ("log_D", "log_alph", "log_lam1", "log_h0", "k_clean", "k_noise"):
        got = g[key].cpu().numpy().reshape(g_o[key].shape) / ms
        fro, mx = rel_err(got, g_o[key])
        assert fro < 2e-4 and mx < 2e-4, (impl, case, key, fro, mx)


def test_fit_reduces_loss_and_checkpoints(tmp_path):
    """model.fit (enhance.py:1152-1157): Adam on the reference's trainable set lowers the masked-MSE loss; weights
    round-trip through save/load; the first Adam step moves parameters along -sign(grad) by ~lr (Keras formula)."""
    from drnmf_b200 import enhance
    F, r, K, B, T = 33, 8, 3, 6, 7
    rng = np.random.default_rng(12)
    W = synth.dictionary(F, 2 * r)
    model = enhance.build_unfolded_snmf(_build_params(F, r, K, T, W))
    clean = (np.abs(rng.standard_normal((B, T, F))) * 1.5).astype(np.float32)
    x = (clean + np.abs(rng.standard_normal((B, T, F))) * 0.8).astype(np.float32)
    x[4, 5:] = -1.0; clean[4, 5:] = -1.0
    w0 = model.get_weights()
    hist = model.fit(x, clean, batch_size=3, epochs=6, validation_data=(x, clean), learning_rate=5e-3,
                     savefile=str(tmp_path / "best.npz"))
    assert len(hist["loss"]) == 6 and hist["loss"][-1] < hist["loss"][0] and hist["val_loss"][-1] < hist["val_loss"][0]
    w1 = model.get_weights()
    names = model.weight_names()
    moved = {n: float(np.abs(a - b).max()) for n, a, b in zip(names, w0, w1)}
    assert moved[[n for n in names if n.endswith("log_D_1")][0]] > 0 and moved["clean_est/kernel"] > 0
    assert moved[[n for n in names if n.endswith("log_U1")][0]] == 0 and moved[[n for n in names if n.endswith("log_lam1")][0]] == 0
    model2 = enhance.build_unfolded_snmf(_build_params(F, r, K, T, W))
    model2.load_weights(str(tmp_path / "best.npz"))
    irm_a = model2.predict_on_batch(x)
    assert irm_a.shape == x.shape and np.isfinite(irm_a).all()


@pytest.mark.parametrize("impl", IMPLS)
def test_ista_ed_golden_and_oracle(golden_dir, impl):
    """enhance.py:402-418 (oracle "A"): golden vectors produced by the reference's own function + a larger case."""
    g = np.load(os.path.join(golden_dir, "ista_ed.npz"))
    cu = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device="cuda")
    H = engine.ista_ed(cu(g["x"]), cu(g["W"]), cu(g["H0"]), float(g["lam1"]), float(g["alph"]), int(g["K"]),
                       impl=None if impl == "tc" else "simt")
    assert max(rel_err(H.cpu().numpy(), g["H"])) < TOL
    rng = np.random.default_rng(4)
    F, R, n, K = 129, 200, 333, 25
    W = synth.dictionary(F, R).astype(np.float64); W /= np.sqrt((W ** 2).sum(0, keepdims=True))
    x = np.abs(rng.standard_normal((F, n))) * 3
    H0 = np.abs(rng.standard_normal((R, n))) * 0.1
    Ho = O.ista_ed(x, W, H0.copy(), 1.0, 60.0, K)
    Hg = engine.ista_ed(cu(x), cu(W), cu(H0), 1.0, 60.0, K, impl=None if impl == "tc" else "simt")
    assert max(rel_err(Hg.cpu().numpy(), Ho)) < TOL


# ---- BASELINE-sized properties (no oracle needed at this size) ----------------------------------------------------
def test_full_size_batch_independence_and_impl_agreement():
    """configs[1] shape (B=64, F=513, R=1000, K=25; T shortened to keep the CUDA-core cross-check quick).
    Utterances are independent: the result for an utterance must not depend on its position in the batch, on the
    batch tile it lands in, or on what else is in the batch; and the tensor-core path must agree with the CUDA-core
    lock path (which is checked against the oracle at small sizes) within the parity tolerance."""
    F, R, K, B, T = 513, 1000, 25, 64, 24
    rng = np.random.default_rng(64)
    p = synth.model_params(F, R, K)
    p["log_U1"], p["log_Uk"] = synth.structured_u_init()
    x = (np.abs(rng.standard_normal((B, T, F))) * 4.0).astype(np.float32)
    lens = rng.integers(T // 2, T + 1, size=B)
    for b in range(B):
        x[b, lens[b]:] = -1.0
    xt = torch.as_tensor(x, device="cuda")
    eng = engine.DrnmfEngine(F, R, K)
    eng.set_params(p)
    H, irm = eng.forward(xt)
    assert eng.recurrent_config()["impl"] == "tcgen05"
    perm = torch.as_tensor(rng.permutation(B), device="cuda")
    Hp, irmp = eng.forward(xt[perm].contiguous())
    assert torch.equal(Hp, H[perm]) and torch.equal(irmp, irm[perm])          # bitwise: fixed reduction order
    Hs, irms = eng.forward(xt[5:18].contiguous())                             # different batch size / tiling
    assert max(rel_err(Hs.cpu().numpy(), H[5:18].cpu().numpy())) < 1e-5
    sim = engine.DrnmfEngine(F, R, K, impl="simt")
    sim.set_params(p)
    H2, irm2 = sim.forward(xt)
    assert max(rel_err(H.cpu().numpy(), H2.cpu().numpy())) < TOL
    assert max(rel_err(irm.cpu().numpy(), irm2.cpu().numpy())) < TOL
    Hn = H.cpu().numpy()
    assert np.isfinite(Hn).all() and (Hn >= 0).all() and (irm > 0).all() and (irm <= 1).all()
    for b in range(0, B, 7):                                                  # masked frames carry the state
        for t in range(lens[b], T):
            assert np.array_equal(Hn[b, t], Hn[b, lens[b] - 1])


def test_snmf_long_contraction_vs_oracle():
    """Frames-long contractions (V H^T, L H^T over n = 4096 frames: split-K + multi-accumulator path) against float64."""
    F, n, R, iters = 129, 4096, 200, 8
    rng = np.random.default_rng(7)
    V = (np.abs(rng.standard_normal((F, n))) * 2).astype(np.float32)
    W0 = (np.abs(rng.standard_normal((F, R))) + 0.1).astype(np.float32)
    H0 = (np.abs(rng.standard_normal((R, n))) + 0.1).astype(np.float32)
    prm = {"cf": "ed", "sparsity": 1.0, "max_iter": iters, "conv_eps": 0.0, "r": R, "init_w": W0, "init_h": H0}
    Wo, Ho, obj = O.sparse_nmf_ed(V, prm, dtype=np.float64)
    Vd, Wd, Hd = (torch.as_tensor(a, device="cuda") for a in (V, W0.copy(), H0.copy()))
    cost, div = engine.snmf_mu_ed(Vd, Wd, Hd, 1.0, iters, 0.0)
    assert max(rel_err(Wd.cpu().numpy(), Wo)) < TOL and max(rel_err(Hd.cpu().numpy(), Ho)) < TOL
    np.testing.assert_allclose(cost, obj["cost"], rtol=2e-5)


@pytest.mark.gpu
def test_snmf_frame_sharded():
    """MU (ED, and KL with the MIN all-reduce of sparse_nmf_gpu.m:201-205) with the frames sharded over 2 ranks
    (all-reduce of V H^T, Lambda H^T and the cost) equals the single-GPU solve.  Two devices: NCCL, one rank per GPU.
    On a 1-GPU box both ranks share the device and reduce over gloo - the same library path and callbacks."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(here, "dist_snmf_check.py")]
    env = dict(os.environ)
    if torch.cuda.device_count() < 2:
        env["DIST_ONE_GPU"] = "1"
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "replicas identical: True" in p.stdout


@pytest.mark.gpu
def test_snmf_inference_irm_and_cache(tmp_path):
    """two-stage dictionary training with the reference's cache naming, inference of the activations of the
    frozen dictionary on new frames (enhance.py:836-845) and the SNMF ratio mask (:847-852) against the oracle."""
    from drnmf_b200 import snmf
    rng = np.random.default_rng(11)
    F, r, n = 65, 12, 300
    clean = np.abs(rng.standard_normal((F, n))).astype(np.float32)
    noisy = clean + np.abs(rng.standard_normal((F, n))).astype(np.float32)
    prm = {"cf": "ed", "sparsity": 0.5, "max_iter": 20.0, "conv_eps": 0.0, "display": 0.0, "random_seed": 2016.0, "r": r,
           "init_w": (np.abs(rng.standard_normal((F, r))) + 0.1).astype(np.float32), "init_h": "ones"}
    noise_init = rng.random((F, r)).astype(np.float32)
    d = str(tmp_path) + "/"
    W1, H1, obj1 = snmf.train_snmf(clean, noisy, prm, noise_init=noise_init, path_dicts=d)
    names = sorted(os.listdir(d))
    assert len(names) == 2 and names[0].startswith("W_clean_") and names[1].startswith("W_noisy_") and names[1].endswith("_sparsity0.500.npz")
    W2, H2, obj2 = snmf.train_snmf(clean, noisy, prm, noise_init=noise_init + 1.0, path_dicts=d)     # served from the cache
    np.testing.assert_array_equal(W1, W2)
    Wo, Ho, _ = O.train_snmf(clean, noisy, prm, noise_init=noise_init)
    assert rel_err(W1, Wo)[0] < 1e-4
    x_new = np.abs(rng.standard_normal((F, 90))).astype(np.float32) + 0.05
    H, obj = snmf.snmf_infer(x_new, W1, prm, max_iter=30)
    po = dict(prm); po.update({"r": 2 * r, "init_w": W1.astype(np.float64), "w_update_ind": np.zeros(2 * r, bool), "conv_eps": 0.0,
                               "max_iter": 30.0})
    _, Href, oref = O.sparse_nmf_ed(x_new, po)
    assert rel_err(H, Href)[0] < 1e-4 and abs(obj["cost"][-1] - oref["cost"][-1]) < 2e-5 * abs(oref["cost"][-1])
    irm = snmf.snmf_irm(W1, H, r)
    assert rel_err(irm, O.snmf_irm(W1.astype(np.float64), Href, r))[0] < 1e-4
    assert irm.shape == x_new.shape and (irm >= 0).all() and (irm <= 1).all()


@pytest.mark.gpu
def test_dataset_padded_tensors_feed_the_network():
    """AudioDataset -> load_data ('mag' features, -1 padding, maxlen chunking) -> the forward pass masks
    exactly the padded frames."""
    from drnmf_b200 import audio_dataset as ad
    N, hop = 128, 32
    waves = [synth.utterance(i, seconds=s)[0] for i, s in enumerate((0.20, 0.11, 0.16))]
    ds = ad.AudioDataset(waves, waves, params_stft={"N": N, "hop": hop, "nch": 1})
    cfg = {"transform_x": "mag", "transform_y": "mag", "maxlen": 40}
    x, y, mask = ad.load_data(cfg, ds)
    T_all = int((ds.fidx[:, 1] - ds.fidx[:, 0]).sum())
    assert x.shape[1] == 40 and int(mask.sum()) == T_all and np.all(x[mask[..., 0] == 0] == -1.0)
    xo, yo, mo = O.reshape_and_pad_stacks(ds.x_stack, ds.y_stack, ds.fidx, O.data_transform("mag"), O.data_transform("mag"), -1.0, 40)
    np.testing.assert_allclose(x, xo, atol=1e-6)
    np.testing.assert_array_equal(mask, mo)
    F, R, K = N // 2 + 1, 32, 3
    p = synth.model_params(F, R, K, alph=40.0)
    eng = engine.DrnmfEngine(F, R, K)
    eng.set_params(p)
    H, irm = eng.forward(torch.as_tensor(x, device="cuda"))
    Ho, irmo = O.drnmf_forward(x, p, dtype=np.float64)
    assert max(rel_err(H.cpu().numpy(), Ho)) < TOL and max(rel_err(irm.cpu().numpy(), irmo)) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("impl", IMPLS)
def test_edge_cases_fully_masked_single_frame_single_utterance(impl):
    """Edge cases of the masked scan: an utterance without any valid frame (outputs stay zero, its state never moves),
    T = 1, B = 1, and a batch whose only valid frame is the last one."""
    F, R, K = 33, 24, 3
    rng = np.random.default_rng(2)
    p = synth.model_params(F, R, K, alph=12.0)
    eng = engine.DrnmfEngine(F, R, K, impl=impl)
    eng.set_params(p)
    # (a) one fully masked utterance next to a normal one
    x = (np.abs(rng.standard_normal((2, 5, F))) * 2).astype(np.float32)
    x[1] = -1.0
    H, irm = eng.forward(torch.as_tensor(x, device="cuda"))
    Ho, irmo = O.drnmf_forward(x, p, dtype=np.float64)
    assert np.all(H[1].cpu().numpy() == 0.0) and np.all(Ho[1] == 0.0)
    assert max(rel_err(H.cpu().numpy(), Ho)) < TOL and max(rel_err(irm[0].cpu().numpy(), irmo[0])) < TOL
    # (b) T = 1, B = 1
    x1 = (np.abs(rng.standard_normal((1, 1, F))) * 2).astype(np.float32)
    H1, irm1 = eng.forward(torch.as_tensor(x1, device="cuda"))
    Ho1, irmo1 = O.drnmf_forward(x1, p, dtype=np.float64)
    assert max(rel_err(H1.cpu().numpy(), Ho1)) < TOL and max(rel_err(irm1.cpu().numpy(), irmo1)) < TOL
    # (c) only the last frame valid: everything before is zero, the last frame starts from h0
    x2 = np.full((3, 4, F), -1.0, np.float32)
    x2[:, -1] = (np.abs(rng.standard_normal((3, F))) * 2).astype(np.float32)
    H2, _ = eng.forward(torch.as_tensor(x2, device="cuda"))
    Ho2 = O.rnn_forward(x2, p)
    assert np.all(H2[:, :-1].cpu().numpy() == 0.0)
    assert max(rel_err(H2.cpu().numpy(), Ho2)) < TOL
    # gradients with a fully masked utterance in the batch: finite, and equal to the batch without it
    y = (x * 0.5).astype(np.float32); y[1] = -1.0
    ls, ms, g = eng.loss_and_grads(torch.as_tensor(x, device="cuda"), torch.as_tensor(y, device="cuda"))
    ls0, ms0, g0 = eng.loss_and_grads(torch.as_tensor(x[:1], device="cuda"), torch.as_tensor(y[:1], device="cuda"))
    assert ms == ms0 == 5.0 and abs(ls - ls0) < 1e-5 * abs(ls0)
    for key in ("log_D", "log_alph", "log_h0", "k_clean"):
        a, b = g[key].cpu().numpy(), g0[key].cpu().numpy()
        assert np.isfinite(a).all() and rel_err(a, b)[0] < 1e-5, key
