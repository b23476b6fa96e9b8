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


class GradBuckets:
    """Per-layer gradient all-reduce for data-parallel training: `layer_ready(k)` starts the all-reduce of
    the k-th bucket of a flat gradient buffer as soon as the backward pass has produced it (on CUDA: ordered after an
    event on the compute stream and issued on a side stream, so that NCCL runs under the weight-gradient GEMMs of the
    later layers); `finish()` reduces the tail (everything after the buckets) together with the loss statistics and
    joins.  Works on CPU tensors over gloo too (no streams) - that is how the logic is tested."""

    def __init__(self, flat, bucket_bounds, tail_start, group=None):
        self.flat, self.bounds, self.tail_start, self.group = flat, list(bucket_bounds), int(tail_start), group
        self.cuda = flat.is_cuda
        self.comm = torch.cuda.Stream(device=flat.device) if self.cuda else None
        self.works = []
        self.launched = 0

    def active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _reduce(self, view):
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ev)
                self.works.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.works.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.launched += 1

    def layer_ready(self, k):
        if self.active() and k < len(self.bounds):
            a, b = self.bounds[k]
            self._reduce(self.flat[a:b])

    def finish(self, loss_sum, mask_sum, reduce_all=False):
        """Returns (global loss_sum, global mask_sum); the flat buffer then holds the gradient sums over all ranks."""
        stats = torch.tensor([loss_sum, mask_sum], dtype=torch.float64, device=self.flat.device)
        if self.active():
            if reduce_all:          # tied parameters: nothing could be started early
                self._reduce(self.flat)
            else:
                self._reduce(self.flat[self.tail_start:])
            self._reduce(stats)
            for w in self.works:
                w.wait()            # on CUDA: the current stream waits for the collective
        self.works = []
        self.launched = 0
        return float(stats[0]), float(stats[1])


class DeviceTrainer:
    """Data-parallel training step that stays on the device (enhance.py:1152-1157 per batch): gradients from
    drnmf_loss_and_grads_cb into ONE flat buffer, per-layer NCCL all-reduce under the backward GEMMs (GradBuckets),
    fused Keras-formula Adam over the flat parameter buffer (drnmf_adam_step) and the rebuild of the derived weights
    (drnmf_set_params).  Parameters use the engine layout: log_D (K|1,F,R), log_alph (K|1, 1|R), log_lam1 (K|1),
    log_h0 (R), k_clean / k_noise (R/2, F)."""

    ORDER = ("log_D", "log_alph", "log_lam1", "log_h0", "k_clean", "k_noise")

    def __init__(self, eng, params, trainable=("log_D", "log_alph", "log_h0", "k_clean", "k_noise"), learning_rate=1e-3,
                 beta_1=0.9, beta_2=0.999, epsilon=1e-8, decay=0.0, group=None, u1=None, uk=None):
        self.eng, self.group = eng, group
        self.lr, self.b1, self.b2, self.eps, self.decay, self.t = learning_rate, beta_1, beta_2, epsilon, decay, 0
        dev = eng.device
        K, F, R = eng.K, eng.F, eng.R
        shapes = {}
        for name in self.ORDER:
            a = params[name]
            a = a.detach().float().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, np.float32)
            if name == "log_D" and a.ndim == 2:
                a = a[None]
            if name == "log_alph":
                a = a.reshape(1, 1) if a.ndim == 0 else (a.reshape(-1, 1) if a.ndim == 1 and a.shape[0] in (1, K) else
                                                          (a.reshape(1, -1) if a.ndim == 1 else a))
            if name == "log_lam1":
                a = a.reshape(-1)
            shapes[name] = a
        offs, off = {}, 0
        for name in self.ORDER:
            offs[name] = off
            off += (shapes[name].size + 3) // 4 * 4          # every segment starts 16-byte aligned
        self.n = off
        self.flat_p = torch.zeros(off, device=dev)
        self.flat_g = torch.zeros(off, device=dev)
        self.flat_m = torch.zeros(off, device=dev)
        self.flat_v = torch.zeros(off, device=dev)
        self.mask = torch.zeros(off, dtype=torch.uint8, device=dev)
        self.p, self.g = {}, {}
        for name in self.ORDER:
            a, o = shapes[name], offs[name]
            self.flat_p[o:o + a.size] = torch.as_tensor(a.reshape(-1), device=dev)
            self.p[name] = self.flat_p[o:o + a.size].view(*a.shape)
            self.g[name] = self.flat_g[o:o + a.size].view(*a.shape) if name != "log_alph" else self.flat_g[o:o + a.size]
            if name in trainable:
                self.mask[o:o + a.size] = 1
        unknown = [t for t in trainable if t not in self.ORDER]
        if unknown:
            raise NotImplementedError("no gradient is implemented for trainable parameter(s) %s" % unknown)
        self.u1 = u1 if u1 is not None else params["log_U1"]
        self.uk = uk if uk is not None else params["log_Uk"]
        nD = shapes["log_D"].shape[0]
        per = F * R
        self.untied_D = nD == K and K > 1
        bounds = [(offs["log_D"] + k * per, offs["log_D"] + (k + 1) * per) for k in range(nD)] if self.untied_D else []
        self.buckets = GradBuckets(self.flat_g, bounds, offs["log_alph"], group)
        self._push_params()

    def _push_params(self):
        d = dict(self.p)
        d["log_U1"], d["log_Uk"] = self.u1, self.uk
        self.eng.set_params(d)

    def train_on_batch(self, x, y, mask_value=-1.0):
        """x, y: (B,T,F) float32 CUDA tensors (or pinned host tensors, copied here).  Returns the global mean loss."""
        dev = self.eng.device
        if not x.is_cuda:
            x, y = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
        ls, ms, _ = self.eng.loss_and_grads(x, y, mask_value, out=self.g,
                                            layer_ready=self.buckets.layer_ready if self.buckets.active() else None)
        ls, ms = self.buckets.finish(ls, ms, reduce_all=not self.untied_D)
        lr = self.lr * (1.0 / (1.0 + self.decay * self.t)) if self.decay > 0 else self.lr
        self.t += 1
        lr_t = lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        self.eng.adam_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, lr_t, self.b1, self.b2, self.eps,
                           grad_scale=1.0 / max(ms, 1.0), trainable=self.mask)
        self._push_params()
        return ls / max(ms, 1.0)


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

    def __init__(self, model, learning_rate=1e-3, clipnorm=0.0, decay=0.0, group=None, loss="mse_of_masked", lam1=0.0):
        self.model, self.group = model, group
        self.loss, self.lam1 = loss, lam1             # 'mse_of_masked' | 'snmf_cost' (enhance.py:1024-1047)
        self.opt = Adam(lr=learning_rate, clipnorm=clipnorm, decay=decay)

    # ---- mapping between the engine's stacked gradients and the model's named tensors -------------------------------
    def _named_grads(self, g):
        rnn = self.model.rnn
        lab = rnn.maps_from_alt.labels_per_k
        out = {}
        have_grad = set()
        for name in ("log_D", "log_alph", "log_lam1"):
            labels = lab[name]
            have_grad.update(labels)
            if len(set(labels)) == 1:
                out[labels[0]] = g[name].reshape(rnn.alt_params[labels[0]].shape)
            else:
                # the engine returns the per-layer gradients stacked (and log_alph flattened to K*alph_dim)
                stacked = g[name].reshape(len(labels), -1)
                for k, l in enumerate(labels):
                    out[l] = stacked[k].reshape(rnn.alt_params[l].shape)
        missing = [k for k in rnn.keys_trainable if k not in have_grad]
        if missing:
            # the reference would train these through Theano autodiff (e.g. log_U1 / log_Uk); the hand-written backward
            # pass has no gradient for them - refuse instead of silently freezing them
            raise NotImplementedError("no gradient is implemented for trainable parameter(s) %s" % missing)
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
        eng.set_training_loss(self.loss, self.lam1)
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
            eng.set_training_loss(self.loss, self.lam1)
            ls, ms, _ = eng.loss_and_grads(xb, yb, self.model.mask_value)
            tot += ls; cnt += ms
        stats = torch.tensor([tot, cnt], dtype=torch.float64, device=eng.device)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(stats, group=self.group)
        return float(stats[0] / stats[1].clamp(min=1.0))

    def fit(self, x, y, batch_size=32, epochs=1, validation_data=None, patience=50, savefile=None, seed=7654, verbose=0):
        """model.fit(x, y, sample_weight=mask, ...) of enhance.py:1152-1157 with ModelCheckpoint(save_best_only) and
        EarlyStopping(monitor='val_loss', patience).  The frame mask is recomputed from the -1 padding (it equals the
        reference's sample_weight).  Returns the history dict {'loss': [...], 'val_loss': [...]}.
        Loss normalisation: L = sum_frames m * mean_F(err^2) / sum_frames m.  In the reference the Masking mask reaches
        the output next to sample_weight, so Keras 2.0.4 divides by mean(mask) once more per batch; that is a per-batch
        constant (exactly 1 for unpadded batches) which rescales the step of ragged batches and the reported loss
        values - this build uses the plain masked mean."""
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
                else:                   # Keras 2.0.4 EarlyStopping.on_epoch_end: test, then count
                    if wait >= patience:
                        break
                    wait += 1
            if verbose:
                print("epoch %d loss %.6f%s" % (ep + 1, hist["loss"][-1],
                                                 " val_loss %.6f" % hist["val_loss"][-1] if hist["val_loss"] else ""))
        return hist
