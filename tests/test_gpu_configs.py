"""GPU parity at the EXACT configurations that are benchmarked (BASELINE.json configs[1..4]), against the float64
oracle / torch.autograd on the float64 oracle -- no transitive argument through a second implementation.

  configs[1]  B=64, T=193, F=513, R=1000, K=25 forward (the plan bench.py runs) vs oracle.drnmf_forward (float64)
  configs[2]  drnmf_loss_and_grads at R=1000, K=25 (MT=8 x KS=8 backward tiling) and at an intermediate multi-tile
              shape (B > 64, leading + fully masked utterances), tied / untied / vector alph, vs torch.autograd
  configs[3]  MU-ED at F=513, R=1000, n=8192 vs oracle.sparse_nmf_ed (float64)
  configs[4]  forward at F=1025 with R=2000 and R=4000 vs the oracle

Tolerances: north_star (1e-4 on H and the mask, Frobenius and max-abs/max); gradients 2e-4.
"""
import os
import sys

import numpy as np
import pytest
import torch

import oracle as O
from drnmf_b200 import engine, synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    fro = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
    mx = np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)
    return fro, mx


def _synthetic_batch(B, T, F, seed, ragged=True):
    rng = np.random.default_rng(seed)
    x = (np.abs(rng.standard_normal((B, T, F))) * 4.0).astype(np.float32)
    lens = np.full(B, T)
    if ragged:
        lens = rng.integers(max(1, T // 2), T + 1, size=B)
        lens[0] = T
        for b in range(B):
            x[b, lens[b]:] = -1.0
    return x, lens


def test_configs1_exact_forward_vs_fp64_oracle():
    """The bench instantiation itself (B=64 x T=193, F=513, R=1000, K=25, structured U) against the float64 oracle."""
    F, R, K, B, T = 513, 1000, 25, 64, 193
    p = synth.model_params(F, R, K)
    x, lens = _synthetic_batch(B, T, F, 193)
    eng = engine.DrnmfEngine(F, R, K)
    pe = dict(p)
    pe["log_U1"], pe["log_Uk"] = synth.structured_u_init()
    eng.set_params(pe)
    H, irm = eng.forward(torch.as_tensor(x, device="cuda"))
    torch.cuda.synchronize()
    cfg = eng.recurrent_config()
    # bench.py runs the library's default plan for this shape and prints it under config.recurrence: same plan here
    assert cfg["impl"] == "tcgen05" and cfg["MT"] == 8, cfg
    assert cfg["groups"] * cfg["n_tiles"] * cfg["NB"] == 64, cfg
    print("configs[1] plan:", cfg)
    Ho, irmo = O.drnmf_forward(x, p, dtype=np.float64)
    fro, mx = rel_err(H.cpu().numpy(), Ho)
    assert fro < TOL and mx < TOL, ("H", fro, mx, cfg)
    fro, mx = rel_err(irm.cpu().numpy(), irmo)
    assert fro < TOL and mx < TOL, ("irm", fro, mx, cfg)


@pytest.mark.parametrize("B", [128, 512])
def test_throughput_mode_forward_vs_fp64_oracle(B):
    """Throughput batches (several batch tiles / batch groups through the same weights), north-star model, short T."""
    F, R, K, T = 513, 1000, 25, 5
    p = synth.model_params(F, R, K)
    x, lens = _synthetic_batch(B, T, F, 500 + B)
    eng = engine.DrnmfEngine(F, R, K)
    eng.set_params(p)
    H, irm = eng.forward(torch.as_tensor(x, device="cuda"))
    cfg = eng.recurrent_config()
    assert cfg["impl"] == "tcgen05", cfg
    Ho, irmo = O.drnmf_forward(x, p, dtype=np.float64)
    assert max(rel_err(H.cpu().numpy(), Ho)) < TOL, (rel_err(H.cpu().numpy(), Ho), cfg)
    assert max(rel_err(irm.cpu().numpy(), irmo)) < TOL, (rel_err(irm.cpu().numpy(), irmo), cfg)


def test_ragged_last_group_forward_vs_fp64_oracle():
    """5 tiles of 64 over 3 batch groups (2, 2, 1): the last group walks the K-slice with its own schedule classes
    (4 sub-chunks per K-slice > 3 TMEM weight buffers, so the tiles of a step reuse what their predecessor left)."""
    F, R, K, T, B = 513, 1000, 6, 3, 320
    p = synth.model_params(F, R, K)
    x, lens = _synthetic_batch(B, T, F, 321)
    eng = engine.DrnmfEngine(F, R, K)
    eng.set_params(p)
    H, irm = eng.forward(torch.as_tensor(x, device="cuda"))
    cfg = eng.recurrent_config()
    assert cfg["impl"] == "tcgen05" and cfg["NB"] == 64 and cfg["n_tiles"] == 2 and cfg["groups"] == 3, cfg
    Ho, irmo = O.drnmf_forward(x, p, dtype=np.float64)
    assert max(rel_err(H.cpu().numpy(), Ho)) < TOL, (rel_err(H.cpu().numpy(), Ho), cfg)


def test_split_epilogue_forward_vs_fp64_oracle():
    """>= 12 tiles of 64 utterances: K-split 2 plan, 4096 outputs per item, warps 4-7 push and own half of the rows
    (recurrent_tc.cu, split epilogue).  Ragged last tile, masked frames, vector alph off / on."""
    F, R, K, T, B = 513, 1000, 5, 3, 800
    for vec in (False, True):
        p = synth.model_params(F, R, K)
        if vec:
            rng = np.random.default_rng(77)
            p["log_alph"] = (p["log_alph"][:, None] + 0.1 * rng.standard_normal((K, R))).astype(np.float32)
        x, lens = _synthetic_batch(B, T, F, 900 + int(vec))
        eng = engine.DrnmfEngine(F, R, K)
        eng.set_params(p)
        H, irm = eng.forward(torch.as_tensor(x, device="cuda"))
        cfg = eng.recurrent_config()
        assert cfg["impl"] == "tcgen05" and cfg["KS"] == 2 and cfg["NB"] == 64, cfg
        Ho, irmo = O.drnmf_forward(x, p, dtype=np.float64)
        assert max(rel_err(H.cpu().numpy(), Ho)) < TOL, (vec, rel_err(H.cpu().numpy(), Ho), cfg)
        assert max(rel_err(irm.cpu().numpy(), irmo)) < TOL, (vec, rel_err(irm.cpu().numpy(), irmo), cfg)


GRAD_CASES = [
    # north-star model: Rp=1024 -> MT=8 x KS=8 clusters, mirrored-block weight loads (scalar alph)
    dict(F=513, R=1000, K=25, B=4, T=6, tied=False, vec=False, alph=None, lead=False),
    # the same tiling with one step size per atom: S_k is not symmetric (sym=0 weight path at scale)
    dict(F=513, R=1000, K=6, B=5, T=5, tied=False, vec=True, alph=None, lead=True),
    # intermediate tiling (Rp=256: MT=2), more than one batch tile, leading-masked and fully masked utterances
    dict(F=129, R=200, K=4, B=70, T=6, tied=False, vec=False, alph=50.0, lead=True),
    dict(F=129, R=200, K=4, B=70, T=6, tied=True, vec=False, alph=50.0, lead=True),
    dict(F=129, R=200, K=3, B=33, T=5, tied=False, vec=True, alph=50.0, lead=True),
    # Rp=512 (MT=4), batch of the training config (32)
    dict(F=257, R=500, K=5, B=32, T=7, tied=False, vec=False, alph=None, lead=True),
]


@pytest.mark.parametrize("case", GRAD_CASES, ids=lambda c: "R%d_K%d_B%d%s%s" % (c["R"], c["K"], c["B"], "_tied" if c["tied"] else "", "_vec" if c["vec"] else ""))
def test_loss_and_grads_at_benchmarked_tilings(case):
    from oracle import torch_oracle as TO
    F, R, K, B, T = (case[k] for k in "FRKBT")
    rng = np.random.default_rng(4242 + R + B)
    p = synth.model_params(F, R, K, alph=case["alph"], lam1=0.5, untied=not case["tied"])
    if not case["tied"]:
        p["log_alph"] = (p["log_alph"] + 0.05 * rng.standard_normal(K)).astype(np.float32)
    if case["vec"]:
        p["log_alph"] = (p["log_alph"][:, None] + 0.1 * rng.standard_normal((K, R))).astype(np.float32)
    x = (np.abs(rng.standard_normal((B, T, F))) * 3).astype(np.float32)
    y = (x * rng.uniform(0.2, 0.9, size=x.shape)).astype(np.float32)
    lens = rng.integers(2, T + 1, size=B)
    lens[0] = T
    for b in range(B):
        x[b, lens[b]:] = -1.0
        y[b, lens[b]:] = -1.0
    n_valid = int(lens.sum())
    if case["lead"]:
        x[1, 0] = -1.0; y[1, 0] = -1.0; n_valid -= 1               # leading masked frame
        n_valid -= int(lens[2]); x[2] = -1.0; y[2] = -1.0          # utterance without any valid frame
    loss_o, g_o, H_o, irm_o = TO.loss_and_grads(x, y, p)
    pe = dict(p)
    if case["tied"]:
        pe["log_D"], pe["log_alph"], pe["log_lam1"] = p["log_D"][:1], p["log_alph"][:1], p["log_lam1"][:1]
        for k in ("log_D", "log_alph", "log_lam1"):
            g_o[k] = g_o[k].sum(axis=0, keepdims=True)
    eng = engine.DrnmfEngine(F, R, K)
    eng.set_params(pe)
    ls, ms, g, irm = eng.loss_and_grads(torch.as_tensor(x, device="cuda"), torch.as_tensor(y, device="cuda"), want_irm=True)
    assert eng.last_backward_impl() == "tcgen05", "the backward chain must run on the persistent tcgen05 kernel"
    assert ms == float(n_valid)
    assert abs(ls / ms - loss_o) < 2e-5 * abs(loss_o), (ls / ms, loss_o)
    valid = (x != -1.0).any(axis=-1)
    assert max(rel_err(irm.cpu().numpy()[valid], irm_o[valid])) < TOL
    for key in ("log_D", "log_alph", "log_lam1", "log_h0", "k_clean", "k_noise"):
        got = g[key].cpu().numpy().reshape(g_o[key].shape) / ms
        fro, mx = rel_err(got, g_o[key])
        assert fro < 2e-4 and mx < 2e-4, (case, key, fro, mx)


@pytest.mark.parametrize("R", [2000, 4000])
def test_configs4_large_dictionary_forward(R):
    """configs[4]: 1025-bin STFT (N_fft 2048), R = 2000 / 4000 atoms, on the persistent tcgen05 kernel."""
    F, K, B, T = 1025, 3, 5, 4
    p = synth.model_params(F, R, K, alph=synth.default_alph(R), lam1=0.5)
    x, lens = _synthetic_batch(B, T, F, R)
    eng = engine.DrnmfEngine(F, R, K)
    eng.set_params(p)
    H, irm = eng.forward(torch.as_tensor(x, device="cuda"))
    cfg = eng.recurrent_config()
    assert cfg["impl"] == "tcgen05", cfg
    Ho, irmo = O.drnmf_forward(x, p, dtype=np.float64)
    assert max(rel_err(H.cpu().numpy(), Ho)) < TOL, (rel_err(H.cpu().numpy(), Ho), cfg)
    assert max(rel_err(irm.cpu().numpy(), irmo)) < TOL, (rel_err(irm.cpu().numpy(), irmo), cfg)


def test_configs3_mu_ed_north_star_dictionary():
    """configs[3] model size: MU-ED at F=513, R=1000 on n=8192 frames (split-K contractions over the frames)."""
    F, n, R, iters = 513, 8192, 1000, 4
    rng = np.random.default_rng(33)
    V = (np.abs(rng.standard_normal((F, n))) * 2).astype(np.float32)
    W0 = (np.abs(rng.standard_normal((F, R))) + 0.1).astype(np.float32)
    H0 = (np.abs(rng.standard_normal((R, n))) + 0.1).astype(np.float32)
    prm = {"cf": "ed", "sparsity": 1.0, "max_iter": iters, "conv_eps": 0.0, "r": R, "init_w": W0, "init_h": H0}
    Wo, Ho, obj = O.sparse_nmf_ed(V, prm, dtype=np.float64)
    Vd, Wd, Hd = (torch.as_tensor(a, device="cuda") for a in (V, W0.copy(), H0.copy()))
    cost, div = engine.snmf_mu_ed(Vd, Wd, Hd, 1.0, iters, 0.0)
    assert max(rel_err(Wd.cpu().numpy(), Wo)) < TOL, rel_err(Wd.cpu().numpy(), Wo)
    assert max(rel_err(Hd.cpu().numpy(), Ho)) < TOL, rel_err(Hd.cpu().numpy(), Ho)
    np.testing.assert_allclose(cost, obj["cost"], rtol=2e-5)
    np.testing.assert_allclose(div, obj["div"], rtol=2e-5)


@pytest.mark.parametrize("cf,beta", [("kl", 1.0), ("is", 0.0), ("beta", 0.5), ("beta", 1.5)])
@pytest.mark.parametrize("impl", ["tc", "simt"])
def test_snmf_beta_divergences_vs_oracle(cf, beta, impl):
    """KL / IS / generic beta-divergence branches of the MU solver (sparse_nmf_gpu.m:201-205, :212-226, :232-260, :266-276)
    with zeros in V, a frozen half of the dictionary on the second run, and the solver's cf/beta selection (:100-115)
    through the Python mirror."""
    from drnmf_b200 import snmf as S
    F, n, R, iters = 129, 700, 96, 5
    rng = np.random.default_rng(77)
    V = (np.abs(rng.standard_normal((F, n))) * 2).astype(np.float32)
    V[rng.random((F, n)) < 0.03] = 0.0
    W0 = (np.abs(rng.standard_normal((F, R))) + 0.1).astype(np.float32)
    H0 = (np.abs(rng.standard_normal((R, n))) + 0.1).astype(np.float32)
    for wu in (None, np.arange(R) >= R // 2):
        prm = {"cf": cf, "beta": beta, "sparsity": 0.4, "max_iter": iters, "conv_eps": 0.0, "r": R, "init_w": W0, "init_h": H0}
        if wu is not None:
            prm["w_update_ind"] = wu
        Wo, Ho, obj = O.sparse_nmf_beta(V, prm, dtype=np.float64)
        W, H, o = S.sparse_nmf_matlab_on_chunk(V, prm, verbose=False, impl=None if impl == "tc" else "simt")
        assert len(o["cost"]) == iters
        assert max(rel_err(W, Wo)) < TOL, (cf, beta, rel_err(W, Wo))
        assert max(rel_err(H, Ho)) < TOL, (cf, beta, rel_err(H, Ho))
        np.testing.assert_allclose(o["div"], obj["div"], rtol=5e-5)
        np.testing.assert_allclose(o["cost"], obj["cost"], rtol=5e-5)
        np.testing.assert_allclose(np.sqrt((W.astype(np.float64) ** 2).sum(0)), 1.0, rtol=1e-5)


@pytest.mark.parametrize("cf", ["ed", "kl"])
def test_snmf_partial_h_and_w_update_masks(cf, monkeypatch):
    """h_update_ind / w_update_ind subsets (sparse_nmf_gpu.m:150-155) through the fused H-update GEMM (row mask in the
    epilogue) and through the two-projection path (DRNMF_MU_UNFUSED=1): same numbers, both equal to the oracle."""
    from drnmf_b200 import snmf as S
    F, n, R, iters = 70, 333, 48, 6
    rng = np.random.default_rng(123)
    V = (np.abs(rng.standard_normal((F, n))) + 0.05).astype(np.float32)
    W0 = (np.abs(rng.standard_normal((F, R))) + 0.1).astype(np.float32)
    H0 = (np.abs(rng.standard_normal((R, n))) + 0.1).astype(np.float32)
    prm = {"cf": cf, "sparsity": 0.3, "max_iter": iters, "conv_eps": 0.0, "r": R, "init_w": W0, "init_h": H0,
           "h_update_ind": np.arange(R) % 3 != 0, "w_update_ind": np.arange(R) < R // 2}
    Wo, Ho, obj = O.sparse_nmf_beta(V, prm, dtype=np.float64)
    outs = []
    for unfused in ("0", "1"):
        if unfused == "1":
            monkeypatch.setenv("DRNMF_MU_UNFUSED", "1")
        W, H, o = S.sparse_nmf_matlab_on_chunk(V, prm, verbose=False)
        assert max(rel_err(W, Wo)) < TOL and max(rel_err(H, Ho)) < TOL, (cf, unfused, rel_err(W, Wo), rel_err(H, Ho))
        np.testing.assert_allclose(o["cost"], obj["cost"], rtol=5e-5)
        frozen = ~prm["h_update_ind"]
        wn = np.sqrt((W0.astype(np.float64) ** 2).sum(0))
        np.testing.assert_allclose(H[frozen], (H0 * wn[:, None])[frozen], rtol=1e-6)      # frozen rows only see the initial rescale
        outs.append((W, H))
    assert max(rel_err(outs[0][1], outs[1][1])) < 1e-5


def test_snmf_beta2_entry_equals_ed_entry():
    """drnmf_snmf_mu_beta(beta = 2) IS the Euclidean solver: bitwise the same iterates as drnmf_snmf_mu_ed."""
    F, n, R = 65, 300, 40
    rng = np.random.default_rng(3)
    V = np.abs(rng.standard_normal((F, n))).astype(np.float32)
    W0 = (np.abs(rng.standard_normal((F, R))) + 0.1).astype(np.float32)
    H0 = (np.abs(rng.standard_normal((R, n))) + 0.1).astype(np.float32)
    outs = []
    for b in (2.0, None):
        Vd, Wd, Hd = (torch.as_tensor(a, device="cuda") for a in (V, W0.copy(), H0.copy()))
        kw = {} if b is None else {"beta": b}
        cost, _ = engine.snmf_mu_ed(Vd, Wd, Hd, 0.3, 6, 0.0, **kw)
        outs.append((Wd.cpu().numpy(), Hd.cpu().numpy(), cost))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])


def test_pipelined_projection_is_bitwise_the_serial_order(monkeypatch):
    """drnmf_forward projects the first frames, starts the persistent recurrence and computes the remaining projections
    next to it (time-major XW, device flag acquired before the first late frame).  Same bits as the serial order, ragged
    batch with a leading-masked and a fully masked utterance, one- and two-tile latency plans."""
    F, R, K = 129, 256, 6
    p = synth.model_params(F, R, K, alph=60.0)
    eng = engine.DrnmfEngine(F, R, K)
    eng.set_params(p)
    for B, T in ((40, 50), (7, 33)):
        x, _ = _synthetic_batch(B, T, F, seed=B)
        x[2, 0] = -1.0
        x[3] = -1.0
        xd = torch.as_tensor(x, device="cuda")
        out = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("DRNMF_FWD_OVERLAP", mode)
            H, irm = eng.forward(xd)
            torch.cuda.synchronize()
            out[mode] = (H.clone(), irm.clone(), eng.stage_times()[1])
        assert torch.equal(out["0"][0], out["1"][0]) and torch.equal(out["0"][1], out["1"][1])
        Ho, irmo = O.drnmf_forward(x, p, dtype=np.float64)
        assert max(rel_err(out["1"][0].cpu().numpy(), Ho)) < TOL and max(rel_err(out["1"][1].cpu().numpy(), irmo)) < TOL


def test_pipelined_projection_falls_back_when_launches_are_serialised():
    """With CUDA_LAUNCH_BLOCKING=1 the second projection chunk cannot run next to the persistent kernel.  The library
    keeps the serial order then; when the pipelined order is forced anyway, the kernel's bounded wait on the projection
    flag expires, the call is redone serially and the handle stays serial (scripts/fwd_overlap_time.py blocking)."""
    import subprocess
    script = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "fwd_overlap_time.py")
    for extra in ({"DRNMF_FWD_OVERLAP": "force"}, {}):
        env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1", **extra)
        r = subprocess.run([sys.executable, script, "blocking"], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "equal to the serial result: True" in r.stdout


def test_pipelined_weight_gradients_are_bitwise_the_serial_order(monkeypatch):
    """drnmf_loss_and_grads launches the late-frame split-K blocks of the weight-gradient GEMMs on a second stream while the
    backward chain still walks the early frames (chain progress word + cuStreamWaitValue32).  Same partial sums, same
    order: the gradients must be bit-identical to the serial order (which the autograd tests pin), ragged batch."""
    F, R, K, B, T = 65, 200, 4, 12, 160            # T * ceil64(B) = 10240 columns -> 8 split-K blocks: the pipelined plan
    p = synth.model_params(F, R, K, alph=60.0)
    eng = engine.DrnmfEngine(F, R, K)
    eng.set_params(p)
    x, _ = _synthetic_batch(B, T, F, seed=5)
    x[2, 0] = -1.0
    xd = torch.as_tensor(x, device="cuda")
    yd = xd * 0.5
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("DRNMF_TRAIN_OVERLAP", mode)
        ls, ms, g = eng.loss_and_grads(xd, yd)
        torch.cuda.synchronize()
        out[mode] = (ls, {k: v.clone() for k, v in g.items()})
    assert out["0"][0] == out["1"][0]
    for k in out["0"][1]:
        assert torch.equal(out["0"][1][k], out["1"][1][k]), k
        assert torch.isfinite(out["1"][1][k]).all()


def test_device_error_is_not_sticky():
    """a latched device-side error word must not poison later calls on the same handle."""
    F, R, K = 33, 16, 2
    p = synth.model_params(F, R, K, alph=10.0)
    eng = engine.DrnmfEngine(F, R, K)
    eng.set_params(p)
    x = torch.rand(2, 3, F, device="cuda")
    eng.forward(x)
    eng.inject_device_error(777)
    with pytest.raises(Exception):
        eng.forward(x)
    H, _ = eng.forward(x)            # the failure was reported once; the handle works again
    assert torch.isfinite(H).all()


def test_device_trainer_step_matches_keras_adam_formula():
    """DeviceTrainer (flat buffers, fused drnmf_adam_step, rebuild) against the Keras-formula Adam in torch applied to the
    engine's own gradients; frozen parameters (log_lam1) must not move; the loss decreases over a few steps."""
    from drnmf_b200 import training
    F, R, K, B, T = 65, 40, 4, 6, 7
    rng = np.random.default_rng(21)
    p = synth.model_params(F, R, K, alph=30.0, lam1=0.5)
    x = (np.abs(rng.standard_normal((B, T, F))) * 2).astype(np.float32)
    y = (x * 0.6).astype(np.float32)
    x[3, 5:] = -1.0; y[3, 5:] = -1.0
    xt, yt = torch.as_tensor(x, device="cuda"), torch.as_tensor(y, device="cuda")
    eng = engine.DrnmfEngine(F, R, K)
    eng.set_params(p)
    ls, ms, g = eng.loss_and_grads(xt, yt)
    ref = training.Adam(lr=2e-3)
    names = ("log_D", "log_alph", "log_h0", "k_clean", "k_noise")
    pt = {n: torch.as_tensor(np.asarray(p[n], np.float32).reshape(g[n].shape), device="cuda").clone() for n in names}
    ref.step(pt, {n: g[n] / ms for n in names})
    eng2 = engine.DrnmfEngine(F, R, K)
    tr = training.DeviceTrainer(eng2, p, learning_rate=2e-3)
    loss0 = tr.train_on_batch(xt, yt)
    assert abs(loss0 - ls / ms) < 1e-6 * abs(loss0)
    for n in names:
        got, want = tr.p[n].reshape(-1), pt[n].reshape(-1)
        assert float((got - want).abs().max()) < 2e-6, (n, float((got - want).abs().max()))
    np.testing.assert_array_equal(tr.p["log_lam1"].cpu().numpy(), np.asarray(p["log_lam1"], np.float32))
    losses = [loss0] + [tr.train_on_batch(xt, yt) for _ in range(5)]
    assert losses[-1] < losses[0]


def test_trainer_untied_vector_alph_config():
    """Trainer/fit with untie_alph and params_untied = [log_D, log_alph] (enhance.py:225-226, 628-633)."""
    from drnmf_b200 import enhance
    F, r, K, B, T = 33, 8, 3, 4, 6
    rng = np.random.default_rng(5)
    W = synth.dictionary(F, 2 * r)
    prm = {"input_dim": F, "hidden_dim": 2 * r, "output_dim": F, "mask_value": -1.0, "maxseq": T, "K_layers": K, "W": W,
           "alph": 20.0, "lam1": 0.5, "params_untied": ["log_D", "log_alph"], "params_trainable": ["log_D", "log_alph"],
           "untie_alph": True}
    model = enhance.build_unfolded_snmf(prm)
    clean = (np.abs(rng.standard_normal((B, T, F))) * 1.5).astype(np.float32)
    x = (clean + np.abs(rng.standard_normal((B, T, F))) * 0.8).astype(np.float32)
    w0 = dict(zip(model.weight_names(), model.get_weights()))
    hist = model.fit(x, clean, batch_size=2, epochs=3, learning_rate=5e-3)
    w1 = dict(zip(model.weight_names(), model.get_weights()))
    assert len(hist["loss"]) == 3 and hist["loss"][-1] < hist["loss"][0]
    a1 = [n for n in w0 if n.endswith("log_alph_1")][0]
    assert w0[a1].shape == (2 * r,) and np.abs(w1[a1] - w0[a1]).max() > 0
    prm2 = dict(prm, params_trainable=["log_D", "log_U1"])
    model2 = enhance.build_unfolded_snmf(prm2)
    with pytest.raises(NotImplementedError):
        model2.fit(x, clean, batch_size=2, epochs=1)


def test_pretrain_snmf_cost_grads_vs_autograd():
    """f4: the optional SNMF pretraining objective (enhance.py:1024-1036) against torch.autograd on the float64 oracle."""
    from oracle import torch_oracle as TO
    F, R, K, B, T = 65, 40, 4, 5, 7
    rng = np.random.default_rng(8)
    p = synth.model_params(F, R, K, alph=30.0, lam1=0.5)
    x = (np.abs(rng.standard_normal((B, T, F))) * 2).astype(np.float32)
    x[2, 4:] = -1.0
    lam1 = 0.7
    loss_o, g_o, _, _ = TO.loss_and_grads(x, x, p, loss="snmf_cost", lam1=lam1)
    eng = engine.DrnmfEngine(F, R, K)
    eng.set_params(p)
    eng.set_training_loss("snmf_cost", lam1)
    xt = torch.as_tensor(x, device="cuda")
    ls, ms, g = eng.loss_and_grads(xt, xt)
    assert abs(ls / ms - loss_o) < 2e-5 * abs(loss_o), (ls / ms, loss_o)
    for key in ("log_D", "log_alph", "log_lam1", "log_h0", "k_clean", "k_noise"):
        got = g[key].cpu().numpy().reshape(g_o[key].shape) / ms
        fro, mx = rel_err(got, g_o[key])
        assert fro < 2e-4 and mx < 2e-4, (key, fro, mx)
    eng.set_training_loss("mse_of_masked")
    with pytest.raises(ValueError):
        eng.set_training_loss("kl")


def test_flag_return_all_hidden_vs_oracle():
    """f4: SimpleDeepRNN(flag_return_all_hidden=True) (custom_layers.py:371-374): all K hidden vectors per frame."""
    from oracle import torch_oracle as TO
    from drnmf_b200 import enhance
    F, R, K, B, T = 40, 24, 3, 4, 6
    rng = np.random.default_rng(3)
    p = synth.model_params(F, R, K, alph=15.0)
    x = (np.abs(rng.standard_normal((B, T, F))) * 2).astype(np.float32)
    x[1, 3:] = -1.0
    x[2, 0] = -1.0
    Hall_o = TO.forward_loss(x, x, p, return_all_hidden=True)[4].detach().numpy()
    eng = engine.DrnmfEngine(F, R, K)
    eng.set_params(p)
    Hall = eng.forward_all_hidden(torch.as_tensor(x, device="cuda")).cpu().numpy()
    assert Hall.shape == (B, T, K * R)
    assert max(rel_err(Hall, Hall_o)) < TOL
    np.testing.assert_array_equal(Hall[1, 3:], np.broadcast_to(Hall[1, 2], (T - 3, K * R)))
    assert np.all(Hall[2, 0] == 0)
    prm = {"input_dim": F, "hidden_dim": R, "output_dim": F, "mask_value": -1.0, "maxseq": T, "K_layers": K, "W": synth.dictionary(F, R),
           "alph": 15.0, "lam1": 0.5, "params_untied": ["log_D", "log_alph"], "params_trainable": ["log_D", "log_alph"],
           "flag_return_all_hidden": True}
    model = enhance.build_unfolded_snmf(prm)
    irm, Hm = model.predict_on_batch(x, return_hidden=True)
    assert Hm.shape == (B, T, K * R) and model.rnn.compute_output_shape((B, T, F)) == (B, T, K * R)
    hist = model.fit_pretrain(x, lam1=0.5, batch_size=2, epochs=2, learning_rate=2e-3)
    assert len(hist["loss"]) == 2 and np.isfinite(hist["loss"]).all()


def test_dataset_from_taskfiles_with_cache(tmp_path):
    """f2: AudioDataset(taskfile_input, taskfile_output, datafile=...) (audio_dataset.py:177-262): wav files listed in
    text files, int16 convention of util.wavread, stacks cached on disk and served from the cache on the second load;
    inputs longer than their targets are clipped (clip_x_to_y)."""
    from drnmf_b200 import audio_dataset as ad, util
    N, hop = 128, 32
    xs, ys = [], []
    for i, s in enumerate((0.10, 0.07)):
        noisy, clean = synth.utterance(i, seconds=s)
        fx, fy = str(tmp_path / ("x%d.wav" % i)), str(tmp_path / ("y%d.wav" % i))
        util.wavwrite(fx, 16000, np.concatenate([noisy, np.zeros(40 * i, np.float32)]).reshape(1, -1))   # second input is longer
        util.wavwrite(fy, 16000, clean.reshape(1, -1))
        xs.append(fx); ys.append(fy)
    tx, ty = tmp_path / "in.txt", tmp_path / "out.txt"
    tx.write_text("\n".join(xs) + "\n"); ty.write_text("\n".join(ys) + "\n")
    cache = str(tmp_path / "data_train")
    ds = ad.AudioDataset(str(tx), str(ty), datafile=cache, params_stft={"N": N, "hop": hop, "nch": 1})
    assert os.path.isfile(cache + ".npz") and ds.x_stack.shape == ds.y_stack.shape and ds.x_stack.shape[0] == N + 2
    win = O.sqrt_hann(N)
    ref = O.stack_reim(O.stft_mc(util.wavread(ys[0]), N, hop, win))
    np.testing.assert_allclose(ds.y_stack[:, ds.fidx[0, 0]:ds.fidx[0, 1]], ref, atol=3e-5)
    ds2 = ad.AudioDataset(str(tx), str(ty), datafile=cache, params_stft={"N": N, "hop": hop, "nch": 1})    # from the cache
    np.testing.assert_array_equal(ds2.x_stack, ds.x_stack)
    np.testing.assert_array_equal(ds2.fidx, ds.fidx)
    x, y, mask = ad.load_data({"transform_x": "mag", "transform_y": "mag", "maxlen": 30}, ds2)
    assert x.shape == y.shape and int(mask.sum()) == int((ds.fidx[:, 1] - ds.fidx[:, 0]).sum())
    with pytest.raises(ValueError):
        ad.AudioDataset(str(tx), str(ty), datafile=cache, params_stft={"N": 256, "hop": 64, "nch": 1})
