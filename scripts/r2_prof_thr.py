"""One throughput-mode forward (B=2048, T=12) for ncu metric captures of k_recurrent_tc."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drnmf_b200 import engine, synth
F, R, K = 513, 1000, 25
B, T = int(os.environ.get("PROF_B", "2048")), int(os.environ.get("PROF_T", "12"))
p = synth.model_params(F, R, K); p["log_U1"], p["log_Uk"] = synth.structured_u_init()
eng = engine.DrnmfEngine(F, R, K); eng.set_params(p)
x = torch.rand(B, T, F, device="cuda") * 4
for _ in range(2):
    eng.forward(x, want_H=False)
torch.cuda.synchronize()
print(eng.recurrent_config(), eng.stage_times())
