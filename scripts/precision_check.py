"""TC vs SIMT vs float64 oracle at the north-star layer shape (error budget of the 3xTF32 tensor-core path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from drnmf_b200 import engine, synth
F, R, K = 513, 1000, 25
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
p = synth.model_params(F, R, K)
def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b), np.abs(a - b).max() / np.abs(b).max()
for T in (24, 96):
    rng = np.random.default_rng(64)
    x = (np.abs(rng.standard_normal((B, T, F))) * 4.0).astype(np.float32)
    Ho, irmo = O.drnmf_forward(x, p, dtype=np.float64)
    H32, irm32 = O.drnmf_forward(x, p, dtype=np.float32)
    print("T=%d oracle fp32 vs fp64: H %.2e/%.2e  irm %.2e/%.2e" % ((T,) + rel(H32, Ho) + rel(irm32, irmo)), flush=True)
    for impl in ("tc", "simt"):
        eng = engine.DrnmfEngine(F, R, K, impl=None if impl == "tc" else "simt")
        eng.set_params(p)
        H, irm = eng.forward(torch.as_tensor(x, device="cuda"))
        H, irm = H.cpu().numpy(), irm.cpu().numpy()
        print("T=%d %-4s vs fp64: H %.2e/%.2e  irm %.2e/%.2e   |H|max %.3f nnz %.3f" % ((T, impl) + rel(H, Ho) + rel(irm, irmo) + (np.abs(Ho).max(), (Ho > 0).mean())), flush=True)
        if impl == "tc":
            for k in (1, 12, 24):
                Wk, Sk, bk = O.layer_weights(p, k)
                ST = eng.derived(0, k).cpu().numpy()[:R, :R]
                print("   S_%d err %.2e/%.2e" % ((k,) + rel(ST, Sk.T)))
