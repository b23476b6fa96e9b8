"""Robustness sweep on the GPU: reference-native shapes, odd batch sizes, long sequences, the SIMT fallback for R > 1024.
Each case is checked against the CUDA-core lock path (itself oracle-checked at small sizes) or the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from drnmf_b200 import engine, synth

def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp(min=1e-300)), float((a - b).abs().max() / b.abs().max().clamp(min=1e-300))

cases = [
    dict(F=257, R=200, K=5, B=250, T=500, alph=50.0),     # reference config r=100: predict slab of 250 x maxlen 500
    dict(F=257, R=200, K=2, B=33, T=77, alph=50.0),
    dict(F=257, R=2000, K=5, B=8, T=20, alph=400.0),      # reference flagship r=1000 -> falls back to the SIMT recurrence
    dict(F=513, R=1000, K=25, B=1, T=193, alph=200.0),    # config[0]: one utterance
    dict(F=1025, R=1000, K=3, B=20, T=15, alph=200.0),    # 2048-point STFT bins
    dict(F=129, R=512, K=4, B=17, T=9, alph=100.0),       # Rp = 512 -> KS = 16? / 8
    dict(F=65, R=384, K=3, B=70, T=5, alph=80.0),         # Rp = 384 -> K-slices straddle M-tiles
    dict(F=40, R=100, K=3, B=300, T=3, alph=30.0),        # many tiles
]
for c in cases:
    F, R, K, B, T = (c[k] for k in "FRKBT")
    rng = np.random.default_rng(F + R)
    p = synth.model_params(F, R, K, alph=c["alph"])
    p["log_U1"], p["log_Uk"] = synth.structured_u_init()
    x = (torch.rand(B, T, F, device="cuda") * 4)
    lens = rng.integers(1, T + 1, size=B); lens[0] = T
    for b in range(B):
        x[b, lens[b]:] = -1.0
    t0 = time.perf_counter()
    eng = engine.DrnmfEngine(F, R, K); eng.set_params(p)
    H, irm = eng.forward(x); torch.cuda.synchronize()
    t1 = time.perf_counter()
    cfg = eng.recurrent_config()
    sim = engine.DrnmfEngine(F, R, K, impl="simt"); sim.set_params(p)
    H2, irm2 = sim.forward(x); torch.cuda.synchronize()
    eh, em = rel(H, H2), rel(irm, irm2)
    ok = max(eh) < 1e-4 and max(em) < 1e-4 and bool(torch.isfinite(H).all())
    print("%s  %s  impl=%s KS=%s NB=%s tiles=%s  H err %.1e/%.1e  irm err %.1e/%.1e  %.0f ms  rec %.2f ms" %
          ("OK  " if ok else "FAIL", c, cfg["impl"], cfg["KS"], cfg["NB"], cfg["n_tiles"], eh[0], eh[1], em[0], em[1],
           1e3 * (t1 - t0), eng.stage_times()[2]), flush=True)
    del eng, sim
    torch.cuda.empty_cache()
# small oracle check of an odd shape through the tc path
F, R, K, B, T = 33, 130, 3, 5, 6
p = synth.model_params(F, R, K, alph=40.0)
x = np.abs(np.random.default_rng(0).standard_normal((B, T, F))).astype(np.float32) * 3
Ho, irmo = O.drnmf_forward(x, p)
eng = engine.DrnmfEngine(F, R, K); eng.set_params(p)
H, irm = eng.forward(torch.as_tensor(x, device="cuda"))
print("oracle check R=130:", np.abs(H.cpu().numpy() - Ho).max() / np.abs(Ho).max(), eng.recurrent_config())
