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


def _mu_sharded_worker(rank, world, port, q):
    """The frame-sharded MU-ED iteration of drnmf_snmf_mu_ed_dist, restated in numpy over gloo: H-update is local to a
    rank's frames; V H^T, Lambda H^T and (div, sum H) are the only cross-rank sums (snmf.cu)."""
    import numpy as np
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    F, R, n, iters, mu = 33, 12, 40, 8, 0.7
    V = np.abs(rng.standard_normal((F, n))) + 0.01
    W = np.abs(rng.standard_normal((F, R))) + 0.01
    H = np.abs(rng.standard_normal((R, n))) + 0.01
    sl = slice(rank * n // world, (rank + 1) * n // world)
    v, h, flr = V[:, sl], H[:, sl].copy(), 1e-9
    wn = np.sqrt((W ** 2).sum(0)); w = W / wn; h *= wn[:, None]
    lam = np.maximum(w @ h, flr)
    costs = []

    def allsum(a):
        t = torch.from_numpy(np.ascontiguousarray(a)); dist.all_reduce(t); return t.numpy()
    for _ in range(iters):
        h = h * (w.T @ v) / np.maximum(w.T @ lam + mu, flr)
        lam = np.maximum(w @ h, flr)
        VH, LH = allsum(v @ h.T), allsum(lam @ h.T)
        dpw = np.maximum(LH + (VH * w).sum(0, keepdims=True) * w, flr)
        w = w * (VH + (LH * w).sum(0, keepdims=True) * w) / dpw
        wn = np.sqrt((w ** 2).sum(0)); w = w / wn
        lam = np.maximum(w @ h, flr)
        sc = allsum(np.array([((v - lam) ** 2).sum(), mu * h.sum()]))
        costs.append(sc[0] + sc[1])
    q.put((rank, w, h, np.array(costs)))
    dist.barrier()
    dist.destroy_process_group()


def test_mu_frame_sharding_gloo():
    """world_size 2: the sharded iteration reproduces the oracle's unsharded MU-ED (W replicated, H sliced, same cost)."""
    import numpy as np
    import oracle as O
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mu_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(3)
    F, R, n, iters, mu = 33, 12, 40, 8, 0.7
    V = np.abs(rng.standard_normal((F, n))) + 0.01
    W = np.abs(rng.standard_normal((F, R))) + 0.01
    H = np.abs(rng.standard_normal((R, n))) + 0.01
    w, h, info = O.sparse_nmf_ed(V, dict(sparsity=mu, max_iter=iters, conv_eps=0.0, init_w=W, init_h=H, r=R))
    for rank, wr, hr, cr in res:
        np.testing.assert_allclose(wr, w, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(hr, h[:, rank * n // 2:(rank + 1) * n // 2], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(cr, info["cost"], rtol=1e-10)
    np.testing.assert_array_equal(res[0][1], res[1][1])


def _bucket_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # flat gradient: 3 per-layer buckets of 4 floats, then a tail of 5 (scalars, h0, recon kernels)
    flat = torch.arange(17, dtype=torch.float32) * (rank + 1)
    b = training.GradBuckets(flat, [(0, 4), (4, 8), (8, 12)], 12)
    for k in range(3):          # what drnmf_loss_and_grads_cb does layer by layer
        b.layer_ready(k)
    loss_sum, mask_sum = b.finish(2.0 * (rank + 1), 10.0 * (rank + 1))
    flat2 = torch.ones(6) * (rank + 1)          # tied dictionary: one reduction of everything at the end
    b2 = training.GradBuckets(flat2, [], 0)
    b2.layer_ready(0)
    l2, m2 = b2.finish(1.0, 1.0, reduce_all=True)
    q.put((rank, loss_sum, mask_sum, flat.clone(), flat2.clone(), m2))
    dist.barrier()
    dist.destroy_process_group()


def test_per_layer_gradient_buckets_gloo():
    """world_size 2 over gloo: per-layer buckets + tail + statistics give the global sums on every rank."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ls, ms, flat, flat2, m2 in res:
        assert ls == 6.0 and ms == 30.0 and m2 == 2.0
        torch.testing.assert_close(flat, torch.arange(17, dtype=torch.float32) * 3.0)
        torch.testing.assert_close(flat2, torch.full((6,), 3.0))


def test_grad_buckets_single_process_is_a_no_op():
    flat = torch.arange(8, dtype=torch.float32)
    b = training.GradBuckets(flat, [(0, 4)], 4)
    b.layer_ready(0)
    assert b.finish(3.0, 2.0) == (3.0, 2.0) and b.launched == 0
    torch.testing.assert_close(flat, torch.arange(8, dtype=torch.float32))
