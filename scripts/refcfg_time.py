import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from drnmf_b200 import engine, synth
F, R, K, B, T = 257, 200, 5, 250, 500
p = synth.model_params(F, R, K, alph=50.0)
eng = engine.DrnmfEngine(F, R, K); eng.set_params(p)
x = torch.rand(B, T, F, device="cuda") * 4
for rep in range(4):
    H, irm = eng.forward(x, want_H=False); torch.cuda.synchronize()
    print(rep, eng.recurrent_config(), ["%.2f" % v for v in eng.stage_times()])
