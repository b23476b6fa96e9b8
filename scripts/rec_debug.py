"""Per-role wait-time breakdown of the persistent recurrence (DRNMF_REC_DEBUG=1) at the north-star shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drnmf_b200 import engine, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 40
F, R, K = 513, 1000, 25
p = synth.model_params(F, R, K)
p["log_U1"], p["log_Uk"] = synth.structured_u_init()
eng = engine.DrnmfEngine(F, R, K)
eng.set_params(p)
x = torch.rand(B, T, F, device="cuda") * 4
for rep in range(2):
    H, irm = eng.forward(x, want_H=False)
    torch.cuda.synchronize()
st = eng.stage_times()
cfg = eng.recurrent_config()
steps = T * (K - 1)
print("B=%d T=%d cfg=%s rec=%.3f ms -> %.2f us/step %.2f us/item" % (B, T, cfg, st[2], 1e3 * st[2] / steps, 1e3 * st[2] / steps / cfg["n_tiles"]))
