"""CPU tests of the host-side training logic, including the N>1 path over gloo (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from drnmf_b200 import training


def test_shard_utterances_partition():
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [training.shard_utterances(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_adam_matches_keras_formula():
    p = {"w": torch.tensor([1.0, -2.0])}
    g = {"w": torch.tensor([0.5, 0.25])}
    opt = training.Adam(lr=1e-3)
    opt.step(p, g)
    # first step: m = 0.1 g, v = 0.001 g^2, lr_t = lr*sqrt(1-b2)/(1-b1) -> p -= lr * g/(|g| + eps*sqrt(1-b2)) ~ lr*sign(g)
    m, v = 0.1 * g["w"], 0.001 * g["w"] ** 2
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    want = torch.tensor([1.0, -2.0]) - lr_t * m / (v.sqrt() + 1e-8)
    torch.testing.assert_close(p["w"], want)
    opt2 = training.Adam(lr=1e-3, clipnorm=0.1)
    p2 = {"w": torch.zeros(2)}
    opt2.step(p2, {"w": torch.tensor([3.0, 4.0])})      # norm 5 -> scaled to 0.1; Adam's first step is ~lr*sign(g) anyway
    assert torch.all(p2["w"] < 0)


def test_allreduce_grads_single_process():
    g = {"a": torch.ones(3), "b": torch.full((2, 2), 2.0)}
    loss, out = training.allreduce_grads(g, 6.0, 4.0)
    assert loss == 1.5
    torch.testing.assert_close(out["a"], torch.full((3,), 0.25))
    torch.testing.assert_close(out["b"], torch.full((2, 2), 0.5))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank r holds gradients of its own utterances: sums r+1, frame counts 10*(r+1)
    g = {"log_D_0": torch.full((4, 3), float(rank + 1)), "log_h0": torch.arange(3, dtype=torch.float32) * (rank + 1)}
    loss, out = training.allreduce_grads(g, loss_sum=2.0 * (rank + 1), mask_sum=10.0 * (rank + 1))
    q.put((rank, loss, out["log_D_0"].clone(), out["log_h0"].clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_gloo():
    """world_size 2 over gloo: every rank ends with the same globally normalised gradients (1-GPU vs N-GPU equality)."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total_frames = 30.0
    for rank, loss, gD, gh in res:
        assert abs(loss - 6.0 / total_frames) < 1e-12
        torch.testing.assert_close(gD, torch.full((4, 3), 3.0 / total_frames))
        torch.testing.assert_close(gh, torch.arange(3, dtype=torch.float32) * 3.0 / total_frames)
