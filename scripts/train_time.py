"""Time one training step (loss + grads) at the north-star shape; prints the stage breakdown via CUDA events."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drnmf_b200 import engine, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 193
F, R, K = 513, 1000, 25
p = synth.model_params(F, R, K)
p["log_U1"], p["log_Uk"] = synth.structured_u_init()
eng = engine.DrnmfEngine(F, R, K)
eng.set_params(p)
x = torch.rand(B, T, F, device="cuda") * 4
y = x * 0.5
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ls, ms, g = eng.loss_and_grads(x, y)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("B=%d T=%d step %.1f ms  loss %.5f  launches so far %d" % (B, T, dt * 1e3, ls / ms, eng.lib.drnmf_launch_count()))
H, irm = eng.forward(x, want_H=False)
torch.cuda.synchronize()
print("forward stages ms", eng.stage_times())
