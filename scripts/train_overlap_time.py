"""Weight-gradient GEMMs pipelined under the backward chain (drnmf_loss_and_grads): step time and gradients with
DRNMF_TRAIN_OVERLAP=1 against the serial order (=0), at the training batch of the reference (32 utterances)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drnmf_b200 import engine, synth
F, R, K = 513, 1000, 25
p = synth.model_params(F, R, K)
p["log_U1"], p["log_Uk"] = synth.structured_u_init()
eng = engine.DrnmfEngine(F, R, K)
eng.set_params(p)
for B, T in ((32, 193), (32, 499), (16, 193)):
    g = torch.Generator(device="cuda").manual_seed(B + T)
    x = torch.rand(B, T, F, device="cuda", generator=g) * 4
    x[1, T // 2:] = -1.0
    y = x * 0.5
    res = {}
    for mode in ("0", "1", "0", "1"):
        os.environ["DRNMF_TRAIN_OVERLAP"] = mode
        best = None
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ls, ms, gr = eng.loss_and_grads(x, y)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        print("B=%d T=%d overlap=%s  loss+grads %.2f ms  loss %.6f" % (B, T, mode, 1e3 * best, ls / ms), flush=True)
        cur = {k: v.clone() for k, v in gr.items()}
        if mode in res:
            assert all(torch.equal(res[mode][k], cur[k]) for k in cur), "run-to-run difference in mode " + mode
        res[mode] = cur
    worst = max(float((res["0"][k] - res["1"][k]).abs().max() / (res["0"][k].abs().max() + 1e-30)) for k in res["0"])
    print("      max relative difference between the two orders: %.2e (bitwise equal: %s)" % (
        worst, all(torch.equal(res["0"][k], res["1"][k]) for k in res["0"])), flush=True)
    assert worst < 1e-5
os.environ.pop("DRNMF_TRAIN_OVERLAP", None)
