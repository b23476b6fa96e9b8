"""torchrun worker: frame-sharded MU-ED over N ranks must reproduce the single-GPU solve.

Launched by tests/test_gpu_parity.py::test_snmf_frame_sharded (needs >= 2 GPUs) or by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tests/dist_snmf_check.py
Every rank solves its slice of the columns of V with a replica of W; rank 0 also runs the unsharded problem and
compares W (replicated), its slice of H, and the cost trace.  Tolerance: the all-reduce changes the summation order
of V H^T and Lambda H^T only (fp32), so 1e-4 relative like the other MU parity tests."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drnmf_b200.engine as eng  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    # DIST_ONE_GPU=1: all ranks share cuda:0 and reduce over gloo (NCCL refuses two ranks on one device) - lets a 1-GPU box
    # exercise the sharded code path and its callbacks
    if os.environ.get("DIST_ONE_GPU") == "1":
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    F, R, n_per, iters, mu = 257, 100, 700, 12, 3.0
    n = n_per * world
    rng = np.random.default_rng(5)
    V = np.abs(rng.standard_normal((F, n))).astype(np.float32) + 0.01
    W0 = np.abs(rng.standard_normal((F, R))).astype(np.float32) + 0.01
    H0 = np.abs(rng.standard_normal((R, n))).astype(np.float32) + 0.01
    sl = slice(rank * n_per, (rank + 1) * n_per)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    ok = True
    # beta = 2 (ED) and beta = 1 (KL) with zeros in V whose smallest positive entry lives on the LAST rank only: the
    # sharded solve must raise the zeros to the GLOBAL minimum (sparse_nmf_gpu.m:201-205; MIN all-reduce)
    for beta in (2.0, 1.0):
        Vb = V.copy()
        if beta != 2.0:
            Vb[rng.random(Vb.shape) < 0.02] = 0.0
            Vb[3, n - 5] = 1e-4
        Vd, Wd, Hd = dev(Vb[:, sl]), dev(W0), dev(H0[:, sl])
        cost, div = eng.snmf_mu_ed(Vd, Wd, Hd, mu, iters, distributed=True, beta=beta)
        torch.cuda.synchronize()
        if rank == 0:
            Vf, Wf, Hf = dev(Vb), dev(W0), dev(H0)
            cost1, div1 = eng.snmf_mu_ed(Vf, Wf, Hf, mu, iters, beta=beta)
            rel = lambda a, b: float((a - b).norm() / b.norm())
            eW, eH = rel(Wd, Wf), rel(Hd, Hf[:, sl])
            ec = float(np.max(np.abs(cost - cost1) / np.abs(cost1)))
            print("dist_snmf world=%d beta=%g relW=%.2e relH=%.2e relcost=%.2e iters=%d" % (world, beta, eW, eH, ec, len(cost)), flush=True)
            ok = ok and eW < 1e-4 and eH < 1e-4 and ec < 1e-5 and len(cost) == len(cost1)
    # W must be bit-identical on every rank (same reduced sums, same update)
    Wall = [torch.empty_like(Wd) for _ in range(world)]
    dist.all_gather(Wall, Wd)
    same = all(torch.equal(Wall[0], w) for w in Wall)
    if rank == 0:
        print("dist_snmf replicas identical: %s" % same, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if (ok and same) else 1)


if __name__ == "__main__":
    main()
