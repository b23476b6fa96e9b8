"""Data-parallel training step (BASELINE configs[2]) at the north-star shape through the public API: every rank holds
B utterances, step = drnmf_loss_and_grads + NCCL all-reduce of the gradients + Keras-formula Adam + parameter rebuild.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 scripts/train_dp_time.py [B T]
(also runs without torchrun on one GPU).  Prints frames/s over all ranks, max-over-ranks device time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from drnmf_b200 import enhance, synth, training

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 193
F, r, K = 513, 500, 25
W = synth.dictionary(F, 2 * r)
prm = {"input_dim": F, "hidden_dim": 2 * r, "output_dim": F, "mask_value": -1.0, "maxseq": T, "K_layers": K, "W": W,
       "alph": synth.default_alph(2 * r), "lam1": 1.0, "params_untied": ["log_D", "log_alph"],
       "params_trainable": ["log_D", "log_alph"]}
model = enhance.build_unfolded_snmf(prm)
tr = training.Trainer(model, learning_rate=1e-4)
rng = np.random.default_rng(100 + rank)          # every rank its own utterances
x = torch.as_tensor((np.abs(rng.standard_normal((B, T, F))) * 3).astype(np.float32)).cuda()
y = x * 0.6
steps, warm = 5, 2
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(warm + steps):
    if i == warm:
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); e0.record()
    loss = tr.train_on_batch(x, y)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# phase breakdown of one more step (device time between events; includes the host-side launch gaps of each phase)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
eng = model._engine_ready()
ev[0].record()
ls, msum, g = eng.loss_and_grads(x, y, model.mask_value)
ev[1].record()
loss2, grads = training.allreduce_grads(tr._named_grads(g), ls, msum, None)
ev[2].record()
tr.opt.step(tr._named_params(), grads)
model._dirty = True
ev[3].record(); torch.cuda.synchronize()
phases = torch.tensor([ev[i].elapsed_time(ev[i + 1]) for i in range(3)], device="cuda")
if world > 1:
    dist.all_reduce(phases, op=dist.ReduceOp.MAX)
if rank == 0:
    print("train_dp phases (ms, max over ranks): loss_and_grads %.1f | all-reduce + normalise %.1f | Adam %.1f"
          % tuple(phases.tolist()), flush=True)
if rank == 0:
    print("train_dp world=%d B/gpu=%d T=%d: %.1f ms per step (max over ranks) -> %.0f frames/s, loss %.5f"
          % (world, B, T, ms.item(), world * B * T / ms.item() * 1e3, loss), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
