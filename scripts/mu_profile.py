import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from drnmf_b200 import engine
F, R, n = 513, 1000, 22528
rng = np.random.default_rng(0)
V = torch.as_tensor(np.abs(rng.standard_normal((F, n))).astype(np.float32), device="cuda")
W = torch.as_tensor((np.abs(rng.standard_normal((F, R))) + 0.1).astype(np.float32), device="cuda")
H = torch.as_tensor((np.abs(rng.standard_normal((R, n))) + 0.1).astype(np.float32), device="cuda")
engine.snmf_mu_ed(V, W, H, 1.0, 3, 0.0)
torch.cuda.synchronize()
