"""torchrun worker: strong scaling of frame-sharded MU-ED dictionary learning (BASELINE configs[3]: 513 x 225,000
frames = 1 h, R = 1000).  Every rank holds n/N frames and a replica of W; V H^T, Lambda H^T and the cost terms are
all-reduced over NCCL (drnmf_snmf_mu_ed_dist).  Prints one JSON line from rank 0: ms per iteration (device time
between barriers, max over ranks), useful TFLOP/s of the whole job."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drnmf_b200.engine as eng

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
F, R = 513, 1000
n_total = int(os.environ.get("MU_FRAMES", "225000"))
iters = int(os.environ.get("MU_ITERS", "100"))
beta = float(os.environ.get("MU_BETA", "2"))      # 2 = ED (configs[3]), 1 = KL, 0 = IS, else the generic branch
n = (n_total // world + 3) // 4 * 4
g = torch.Generator(device="cuda").manual_seed(11 + rank)
V = torch.rand(F, n, device="cuda", generator=g) * 4
gw = torch.Generator(device="cuda").manual_seed(7)
W = torch.rand(F, R, device="cuda", generator=gw) + 0.1
H = torch.rand(R, n, device="cuda", generator=g) + 0.1
eng.snmf_mu_ed(V, W, H, 1.0, 2, distributed=world > 1, beta=beta)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
cost, div = eng.snmf_mu_ed(V, W, H, 1.0, iters, distributed=world > 1, beta=beta)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item()) / len(cost)
if rank == 0:
    print(json.dumps({"workload": "MU sparse NMF (beta = %g), F=%d, R=%d, %d frames total, %d iterations, explicit inits" % (beta, F, R, n * world, len(cost)),
                      "n_gpus": world, "frames_per_gpu": n, "ms_per_iteration": ms, "ms_per_100_iterations": 100 * ms,
                      "useful_tflops_total": 12.0 * F * R * n * world / (ms / 1e3) / 1e12, "cost_first": float(cost[0]), "cost_last": float(cost[-1]),
                      "collective": "nccl all_reduce of V H^T and Lambda H^T (F x 1024 floats each) + (div, mu*sum H) per iteration"}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
