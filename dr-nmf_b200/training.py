"""Training driver of the unfolded network, mirroring what the reference asks of Keras (enhance.py:1040-1073,
1131-1166): 'mse' of (x * irm) against y with temporal sample weights, Adam(lr, clipnorm, decay), best-only
checkpointing and early stopping.  Gradients come from the CUDA backward pass (drnmf_loss_and_grads); the optimizer
update itself is plain torch arithmetic on the parameter tensors (optimizer/callbacks are not kernels).

Data parallel: every rank runs the same model on its own utterances; one all-reduce(sum) of the flattened gradients
plus (loss_sum, frame_count) per step (NCCL on GPUs, gloo in the CPU tests), then every rank applies the same update.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist


def shard_utterances(n_utt, world_size, rank):
    """Contiguous, near-equal utterance ranges per rank (inference and training shard by utterance only: the time and
    layer dimensions are a serial chain)."""
    base, rem = divmod(int(n_utt), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_grads(grads, loss_sum, mask_sum, group=None):
    """Sum gradients and the loss statistics over ranks with ONE collective, then normalise by the global number of
    valid frames (the loss is a masked mean).  grads: dict name -> tensor.  Returns (loss, grads) with grads of the
    mean loss.  Works on CUDA tensors (NCCL) and CPU tensors (gloo)."""
    names = sorted(grads)
    dev = grads[names[0]].device
    stats = torch.tensor([loss_sum, mask_sum], dtype=torch.float64, device=dev)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        flat = torch.cat([grads[n].reshape(-1).to(torch.float32) for n in names])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
        off = 0
        out = {}
        for n in names:
            k = grads[n].numel()
            out[n] = flat[off:off + k].reshape(grads[n].shape)
            off += k
        grads = out
    total = float(stats[1])
    inv = 1.0 / max(total, 1.0)
    return float(stats[0]) * inv, {n: g * inv for n, g in grads.items()}


class Adam:
    """Keras 2.0.4 Adam: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); p -= lr_t * m / (sqrt(v) + eps); optional
    clipnorm (global l2, 0 = off as in every shipped config) and 1/(1 + decay * iterations) schedule."""

    def __init__(self, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-8, decay=0.0, clipnorm=0.0):
        self.lr, self.b1, self.b2, self.eps, self.decay, self.clipnorm = lr, beta_1, beta_2, epsilon, decay, clipnorm
        self.t = 0
        self.m, self.v = {}, {}

    def step(self, params, grads):
        if self.clipnorm and self.clipnorm > 0:
            norm = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
            if norm > self.clipnorm:
                grads = {k: g * (self.clipnorm / norm) for k, g in grads.items()}
        lr = self.lr * (1.0 / (1.0 + self.decay * self.t)) if self.decay > 0 else self.lr
        self.t += 1
        lr_t = lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k, g in grads.items():
            p = params[k]
            if k not in self.m:
                self.m[k], self.v[k] = torch.zeros_like(p), torch.zeros_like(p)
            self.m[k].mul_(self.b1).add_(g, alpha=1.0 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
            p.sub_(lr_t * self.m[k] / (self.v[k].sqrt() + self.eps))


class Trainer:
    """model: enhance.UnfoldedSNMFModel.  Trainable tensors follow the reference: keys_trainable of the RNN layer
    (log_D_k, log_alph_k in the shipped configs), log_h0, and the two DenseNonNegW kernels (enhance.py:283,292)."""

    def __init__(self, model, learning_rate=1e-3, clipnorm=0.0, decay=0.0, group=None):
        self.model, self.group = model, group
        self.opt = Adam(lr=learning_rate, clipnorm=clipnorm, decay=decay)

    # ---- mapping between the engine's stacked gradients and the model's named tensors -------------------------------
    def _named_grads(self, g):
        rnn = self.model.rnn
        lab = rnn.maps_from_alt.labels_per_k
        out = {}
        for name in ("log_D", "log_alph", "log_lam1"):
            labels = lab[name]
            if len(set(labels)) == 1:
                out[labels[0]] = g[name].reshape(rnn.alt_params[labels[0]].shape) if name != "log_D" else g[name][0]
            else:
                for k, l in enumerate(labels):
                    out[l] = g[name][k].reshape(rnn.alt_params[l].shape)
        out = {k: v for k, v in out.items() if k in rnn.keys_trainable}
        out["log_h0"] = g["log_h0"]
        out["clean_est/kernel"], out["noise_est/kernel"] = g["k_clean"], g["k_noise"]
        return out

    def _named_params(self):
        m = self.model
        dev = m.rnn.log_h0.device
        m.clean_est.kernel = m.clean_est.kernel.to(dev)
        m.noise_est.kernel = m.noise_est.kernel.to(dev)
        p = {k: v for k, v in m.rnn.alt_params.items() if k in m.rnn.keys_trainable}
        p["log_h0"] = m.rnn.log_h0
        p["clean_est/kernel"], p["noise_est/kernel"] = m.clean_est.kernel, m.noise_est.kernel
        return p

    def train_on_batch(self, x, y):
        eng = self.model._engine_ready()
        dev = eng.device
        xt = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(dev) if not torch.is_tensor(x) else x
        yt = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float32)).to(dev) if not torch.is_tensor(y) else y
        ls, ms, g = eng.loss_and_grads(xt, yt, self.model.mask_value)
        loss, grads = allreduce_grads(self._named_grads(g), ls, ms, self.group)
        self.opt.step(self._named_params(), grads)
        self.model._dirty = True            # derived tensors (Gram matrices, ...) are rebuilt before the next forward
        return loss

    def evaluate(self, x, y, batch_size=32):
        eng = self.model._engine_ready()
        tot = cnt = 0.0
        for s in range(0, len(x), batch_size):
            xb = torch.as_tensor(np.ascontiguousarray(x[s:s + batch_size], dtype=np.float32)).to(eng.device)
            yb = torch.as_tensor(np.ascontiguousarray(y[s:s + batch_size], dtype=np.float32)).to(eng.device)
            ls, ms, _ = eng.loss_and_grads(xb, yb, self.model.mask_value)
            tot += ls; cnt += ms
        stats = torch.tensor([tot, cnt], dtype=torch.float64, device=eng.device)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(stats, group=self.group)
        return float(stats[0] / stats[1].clamp(min=1.0))

    def fit(self, x, y, batch_size=32, epochs=1, validation_data=None, patience=50, savefile=None, seed=7654, verbose=0):
        """model.fit(x, y, sample_weight=mask, ...) of enhance.py:1152-1157 with ModelCheckpoint(save_best_only) and
        EarlyStopping(monitor='val_loss', patience).  The frame mask is recomputed from the -1 padding (it equals the
        reference's sample_weight).  Returns the history dict {'loss': [...], 'val_loss': [...]}."""
        rng = np.random.default_rng(seed)
        hist = {"loss": [], "val_loss": []}
        best, wait = float("inf"), 0
        n = len(x)
        for ep in range(epochs):
            perm = rng.permutation(n)
            losses = []
            for s in range(0, n, batch_size):
                idx = np.sort(perm[s:s + batch_size])
                losses.append(self.train_on_batch(x[idx], y[idx]))
            hist["loss"].append(float(np.mean(losses)))
            if validation_data is not None:
                vl = self.evaluate(validation_data[0], validation_data[1], batch_size)
                hist["val_loss"].append(vl)
                if vl < best:
                    best, wait = vl, 0
                    if savefile:
                        self.model.save_weights(savefile)
                else:
                    wait += 1
                    if wait >= patience:
                        break
            if verbose:
                print("epoch %d loss %.6f%s" % (ep + 1, hist["loss"][-1],
                                                 " val_loss %.6f" % hist["val_loss"][-1] if hist["val_loss"] else ""))
        return hist
