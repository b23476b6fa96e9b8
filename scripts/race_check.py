"""Bitwise repeatability of the persistent recurrence at the north-star shape (same input run N times + permuted)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drnmf_b200 import engine, synth
F, R, K, B, T = 513, 1000, 25, int(sys.argv[1]) if len(sys.argv) > 1 else 64, 24
rng = np.random.default_rng(64)
p = synth.model_params(F, R, K)
p["log_U1"], p["log_Uk"] = synth.structured_u_init()
x = (np.abs(rng.standard_normal((B, T, F))) * 4.0).astype(np.float32)
xt = torch.as_tensor(x, device="cuda")
eng = engine.DrnmfEngine(F, R, K)
eng.set_params(p)
H0, _ = eng.forward(xt)
H0 = H0.clone()
bad = 0
for rep in range(8):
    H, _ = eng.forward(xt)
    if not torch.equal(H, H0):
        bad += 1
        d = (H - H0).abs()
        idx = torch.nonzero(d > 0)
        print("rep", rep, "mismatch: n=%d max=%.3e first idx=%s (b,t,r)" % (idx.shape[0], d.max().item(), idx[0].tolist()),
              "rows", sorted(set((idx[:, 2] // 16).tolist()))[:12], "t", sorted(set(idx[:, 1].tolist()))[:6])
print(os.environ.get("TAG", ""), eng.recurrent_config(), "mismatching runs:", bad)
