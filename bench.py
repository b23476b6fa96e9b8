#!/usr/bin/env python
"""bench.py -- DR-NMF frames/s on B200 (driver contract: one JSON line on stdout from rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W]            product arm (CUDA path through libdrnmf.so)
  python bench.py --impl reference [...]                          the reference's algorithm on the host cores
                                                                  (numpy restatement in oracle/: the reference
                                                                  itself is py2+Theano+MATLAB and cannot run)

Workload (BASELINE.json configs[1]): DR-NMF forward as enhance.py drives it -- 25 untied unfolded layers, R=1000
atoms (r=500 speech + 500 noise), 513 bins (N_fft 1024, hop 256), Wiener-style mask, iSTFT -- on a batch of 64
synthetic 3 s / 16 kHz noisy utterances (193 frames each) per GPU.  A "step" is one pass over one such batch.
  value  frames/s with the magnitudes + STFT stack already resident in HBM (forward + mask + iSTFT timed)
  e2e    the same through drnmf_enhance_host: pinned HOST buffers in, enhanced audio on the HOST out, copies timed
Multi-GPU: utterances are independent -> each rank processes its own batch (weak scaling), no collective in the
timed region; barrier + synchronize on both sides, device time, max over ranks.
STFT analysis (waveform -> magnitudes + [Re;Im] stack) happens ONCE before the timed regions, as in the reference's
cached-features flow (audio_dataset.py:194): it is outside both `value` and `e2e` (its kernel is timed in extras).

  python bench.py --workload train [--gpus N]   BASELINE configs[2]: data-parallel training step (B=32 utterances per
      GPU, forward + BPTT + weight gradients, per-layer NCCL all-reduce under the backward GEMMs, fused Adam, rebuild
      of the derived weights), all inside the timed region; run under torch.distributed.run for N > 1.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1; the CPU arms of this benchmark are meant to use every host core the BLAS can
# take, so the thread count is restored before numpy (and its BLAS) is imported.
if "--impl" in sys.argv and "reference" in sys.argv or os.environ.get("OMP_NUM_THREADS") == "1":
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    os.environ.pop("MKL_NUM_THREADS", None)
    os.environ.pop("OPENBLAS_NUM_THREADS", None)

# Exactly ONE line goes to stdout (the JSON).  Everything else that libraries print there (e.g. NCCL's version banner)
# is diverted to stderr by pointing fd 1 at fd 2 for the duration of the run.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DR-NMF frames/s (25 layers, R=1000)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="utterances per GPU (BASELINE configs[1]: 64)")
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--nfft", type=int, default=1024)
    ap.add_argument("--hop", type=int, default=256)
    ap.add_argument("--R", type=int, default=1000)
    ap.add_argument("--layers", type=int, default=25)
    ap.add_argument("--cpu-utts", type=int, default=64, help="utterances in the CPU-baseline sample")
    ap.add_argument("--ref-utts", type=int, default=64, help="utterances per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="infer", choices=["infer", "train"])
    ap.add_argument("--train-batch", type=int, default=32, help="utterances per GPU of --workload train (enhance.py:1153)")
    ap.add_argument("--no-throughput", action="store_true", help="skip the B=512 / B=2048 throughput-mode timings")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the training step, sparse-NMF and STFT timings reported under config.extras")
    ap.add_argument("--no-parity", action="store_true", help="skip the float64-oracle parity check of one untimed step")
    ap.add_argument("--mu-frames", type=int, default=225000, help="frames of the full-size MU-ED timing (configs[3]: 1 h)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16": p["bf16_tflops"], "bf16_sus": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "src": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16": 1590.0, "bf16_sus": 1400.0, "src": "fallback"}


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.th = index, [], False, None

    def _loop(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.15)

    def start(self):
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_workload(args, rank):
    """Synthetic noisy utterances of this rank: waveforms, dictionary parameters."""
    from drnmf_b200 import synth
    B = args.batch
    pairs = [synth.utterance(rank * B + i, seconds=args.seconds) for i in range(B)]
    F = args.nfft // 2 + 1
    p = synth.model_params(F, args.R, args.layers, alph=synth.default_alph(args.R), lam1=1.0)
    p["log_U1"], p["log_Uk"] = synth.structured_u_init()     # (diag, off) of build_alt's U_1 / U_k, no R x R arrays
    return pairs, p


def _lib_check_istft(engine, plan, stack):
    from drnmf_b200 import _lib
    _lib.check(plan.eng.lib.drnmf_mask_istft(engine._ptr(stack), engine._ptr(plan.irm), engine._ptr(plan.fidx), engine._ptr(plan.out_offs),
                                             plan.B, plan.max_frames, plan.N, plan.hop, plan.B * plan.T, engine._ptr(plan.audio),
                                             plan.ws_ptr, plan.nb, engine._stream()))


def flops_per_frame(F, R, K):
    """2FR(K+1) + 2R^2(K-1) (Gram form, as the reference computes it); recurrence part separately."""
    return 2.0 * F * R * (K + 1) + 2.0 * R * R * (K - 1), 2.0 * R * R * (K - 1)


# ------------------------------------------------------------------------------------------------
def cpu_forward_frames_per_s(pairs, p, args, n_utts, reps=1):
    """The oracle (float32, all host threads through the BLAS) on a bounded sample: STFT magnitudes -> DR-NMF
    forward -> mask -> iSTFT, exactly what the GPU arm computes."""
    import oracle as O
    N, hop = args.nfft, args.hop
    F = N // 2 + 1
    win = O.sqrt_hann(N)
    pp = dict(p)
    R = args.R
    if isinstance(pp["log_U1"], tuple):
        e = np.float32(1e-7)
        pp["log_U1"] = np.log(e + np.eye(R, dtype=np.float32))
        pp["log_Uk"] = np.log(e + np.zeros((R, R), dtype=np.float32))
    sample = pairs[:n_utts]
    stacks = [O.stack_reim(O.stft_mc(n, N, hop, win)) for n, _ in sample]
    T = stacks[0].shape[1]
    x = np.stack([O.magnitude(s).T for s in stacks]).astype(np.float32)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        H, irm = O.drnmf_forward(x, pp, dtype=np.float32)
        for b, s in enumerate(stacks):
            O.reconstruct_x(s, hop, win, mask=irm[b].T.astype(np.float32))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return len(sample) * T / best, best, T


def run_reference(args, rank, world):
    """--impl reference: rank 0 alone times the CPU restatement of the reference path."""
    if rank != 0:
        return
    try:
        from threadpoolctl import threadpool_info
        cores = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:
        cores = os.cpu_count() or 1
    pairs, p = build_workload(argparse.Namespace(**{**vars(args), "batch": args.ref_utts}), 0)
    if args.warmup > 0:
        cpu_forward_frames_per_s(pairs, p, args, min(2, args.ref_utts))
    times = []
    fps_T = None
    for _ in range(args.steps):
        fps, dt, T = cpu_forward_frames_per_s(pairs, p, args, args.ref_utts)
        times.append(dt); fps_T = T
    ms = 1e3 * float(np.mean(times))
    value = args.ref_utts * fps_T / (ms / 1e3)
    sample = "%d of the %d utterances of one batch per step (3 s each, %d frames)" % (args.ref_utts, args.batch, fps_T)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DR-NMF forward + mask + iSTFT, numpy restatement of the reference (oracle/) on host cores",
                   "utterances_per_step": args.ref_utts, "T": fps_T, "F": args.nfft // 2 + 1, "R": args.R,
                   "K_layers": args.layers, "N_fft": args.nfft, "hop": args.hop},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
TRAIN_METRIC = "DR-NMF training frames/s (25 layers, R=1000, data-parallel)"


def run_train_reference(args, rank, world):
    """--impl reference --workload train: the float64 torch restatement of the reference's forward + loss with
    torch.autograd (the reference trains through Theano autodiff) on a bounded sample, rank 0 only."""
    if rank != 0:
        return
    import torch
    from oracle import torch_oracle as TO
    from drnmf_b200 import synth
    F = args.nfft // 2 + 1
    R, K = args.R, args.layers
    Bs, Ts = 2, 16
    p = synth.model_params(F, R, K, alph=synth.default_alph(R), lam1=1.0)
    rng = np.random.default_rng(1)
    x = (np.abs(rng.standard_normal((Bs, Ts, F))) * 3).astype(np.float32)
    y = (x * 0.5).astype(np.float32)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        TO.loss_and_grads(x, y, p)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    value = Bs * Ts / (ms / 1e3)
    sample = "%d utterances x %d frames per step (float64 torch autograd restatement, gradient only, no optimizer)" % (Bs, Ts)
    emit({"impl": "reference", "metric": TRAIN_METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
          "data": "synthetic", "config": {"workload": "configs[2] training step, CPU restatement on a bounded sample", "F": F, "R": R, "K_layers": K},
          "cpu_baseline": {"value": value, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
          "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


def run_train(args, rank, world, local_rank):
    """BASELINE configs[2]: one data-parallel training step per 'step' (enhance.py:1152-1157 per batch): every rank
    holds --train-batch synthetic utterances (noisy magnitudes in, clean magnitudes as target), gradients are
    all-reduced per layer over NCCL under the backward GEMMs, Adam + rebuild of the derived weights run on the device.
    Everything is inside the timed region; the e2e arm also copies the batch from pinned host memory every step."""
    import torch
    import torch.distributed as dist
    from drnmf_b200 import engine, training, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the DR-NMF path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    N, hop, B, R, K = args.nfft, args.hop, args.train_batch, args.R, args.layers
    F = N // 2 + 1
    pairs = [synth.utterance(rank * B + i, seconds=args.seconds) for i in range(B)]
    lens = [len(n) for n, _ in pairs]
    offs = np.concatenate([[0], np.cumsum(lens)])[:-1]
    mags = []
    for which in (0, 1):
        audio = torch.as_tensor(np.concatenate([pr[which] for pr in pairs]), device=dev)
        _, mag, fidx = engine.stft_mag(audio, list(offs), lens, N, hop, want_stack=False)
        mags.append(mag)
    T = int(fidx[0, 1] - fidx[0, 0])
    x_dev, y_dev = mags[0].reshape(B, T, F).contiguous(), mags[1].reshape(B, T, F).contiguous()
    x_host, y_host = x_dev.cpu().pin_memory(), y_dev.cpu().pin_memory()
    p = synth.model_params(F, R, K, alph=synth.default_alph(R), lam1=1.0)
    p["log_U1"], p["log_Uk"] = synth.structured_u_init()
    eng = engine.DrnmfEngine(F, R, K)
    tr = training.DeviceTrainer(eng, p, learning_rate=1e-4)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        l0 = eng.lib.drnmf_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            loss = fn()
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), eng.lib.drnmf_launch_count() - l0, loss

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, launches, loss_a = timed(lambda: tr.train_on_batch(x_dev, y_dev), args.steps, args.warmup)
    ms_e2e, _, loss_b = timed(lambda: tr.train_on_batch(x_host, y_host), args.steps, args.warmup)
    clocks = sampler.stop()
    frames = world * B * T * args.steps
    fl_fwd, _ = flops_per_frame(F, R, K)
    pk = peaks()
    tf32_peak = pk["bf16_sus"] / 2.0
    achieved = 3.0 * fl_fwd * B * T / (ms_dev / args.steps / 1e3) / 1e12          # forward + ~2x for the backward contractions
    n_grad = tr.n
    if rank == 0:
        emit({"metric": TRAIN_METRIC, "value": frames / (ms_dev / 1e3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
              "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "f32 (tcgen05 kind::tf32 x3 error-compensated products, fp32 accumulate)", "data": "synthetic",
              "config": {"workload": "configs[2]: DR-NMF training step (BPTT through %d unfolded layers for W_k / alph_k), "
                                     "%d synthetic utterances per GPU, data-parallel with per-layer NCCL gradient all-reduce" % (K, B),
                         "B_per_gpu": B, "T": T, "F": F, "R": R, "K_layers": K, "optimizer": "Adam (Keras 2.0.4 formula), fused kernel",
                         "forward": eng.recurrent_config(), "backward": eng.recurrent_config(backward=True),
                         "collective": {"kind": "nccl all_reduce(sum), %d per-layer buckets + 1 tail + 1 statistics, started from the "
                                                "library's layer-ready callback" % (K if tr.untied_D else 0),
                                        "bytes_per_step": int(n_grad * 4), "inside_timed_region": True},
                         "loss_first_timed_arm": float(loss_a), "loss_e2e_arm": float(loss_b),
                         "l2": "activations + deltas (%.1f GB) exceed the 126 MB L2; no explicit flush" % (4.0 * K * eng.Rp * T * 64 * 4 / 1e9)},
              "clocks": clocks,
              "e2e": {"value": frames / (ms_e2e / 1e3), "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                      "h2d_bytes_per_step": int(2 * x_host.numel() * 4), "d2h_bytes_per_step": 16},
              "gpu_launches": int(launches),
              "roofline": {"bound": "tensor", "kernel": "whole training step (projection, recurrence fwd + bwd, weight-gradient GEMMs)",
                           "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
                           "pipe_frac_3xtf32": 3.0 * achieved / tf32_peak, "traffic": None,
                           "peak_source": "%s MEASURED_PEAKS.json bf16_tflops_sustained/2 (TF32 dense)" % pk["src"],
                           "note": "algorithmic FLOPs = 3 x forward (forward + two backward contractions per product)"},
              "cpu_baseline": None})
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.workload == "train":
            run_train_reference(args, rank, world)
        else:
            run_reference(args, rank, world)
        return
    if args.workload == "train":
        run_train(args, rank, world, local_rank)
        return

    import torch
    import torch.distributed as dist
    from drnmf_b200 import engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the DR-NMF path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    N, hop, B, R, K = args.nfft, args.hop, args.batch, args.R, args.layers
    F = N // 2 + 1

    pairs, p = build_workload(args, rank)
    lens = [len(n) for n, _ in pairs]
    offs = np.concatenate([[0], np.cumsum(lens)])[:-1]
    audio = torch.as_tensor(np.concatenate([n for n, _ in pairs]), device=dev)
    stack, mag, fidx = engine.stft_mag(audio, list(offs), lens, N, hop)          # inputs are made by the CUDA STFT
    T = int(fidx[0, 1] - fidx[0, 0])
    x_dev = mag.reshape(B, T, F).contiguous()
    frames = np.full((B,), T, np.int32)

    eng = engine.DrnmfEngine(F, R, K)
    eng.set_params(p)
    plan = engine.EnhancePlan(eng, B, T, N, hop)
    torch.cuda.synchronize()

    # host copies for the end-to-end arm (pinned)
    x_host = x_dev.cpu().pin_memory()
    stack_host = stack.cpu().pin_memory()
    frames_host = torch.as_tensor(frames)
    out_host = torch.empty((B, plan.L), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, with_stage=False):
        for _ in range(warmup):
            fn()
        barrier()
        launches0 = eng.lib.drnmf_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage = np.zeros(4)
        ev0.record()
        for _ in range(steps):
            fn()
            if with_stage:
                stage += np.array(eng.stage_times())      # CUDA events recorded by the library on this stream
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), eng.lib.drnmf_launch_count() - launches0, stage / max(steps, 1)

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, launches, stage_ms = timed(lambda: plan.run(x_dev, stack), args.steps, args.warmup, with_stage=True)
    ms_e2e, _, _ = timed(lambda: eng.enhance_host(x_host, stack_host, frames_host, N, hop, out=out_host),
                         args.steps, args.warmup)
    clocks = sampler.stop()

    frames_total = world * B * T * args.steps
    value = frames_total / (ms_dev / 1e3)
    e2e_value = frames_total / (ms_e2e / 1e3)
    fl_all, fl_rec = flops_per_frame(F, R, K)
    pk = peaks()
    tf32_peak = pk["bf16_sus"] / 2.0               # TF32 dense = half the measured sustained bf16 rate (kernel inside a long step)
    rec_cfg = eng.recurrent_config()
    rec_ms = float(stage_ms[2])
    achieved = fl_rec * B * T / (rec_ms / 1e3) / 1e12 if rec_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "k_recurrent_tc (persistent recurrence over T x K_layers)" if rec_cfg["impl"] == "tcgen05"
                else "k_step_simt", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": achieved / tf32_peak, "pipe_frac_3xtf32": 3.0 * achieved / tf32_peak, "traffic": None,
                "peak_source": "%s MEASURED_PEAKS.json bf16_tflops_sustained/2 (TF32 dense)" % pk["src"],
                "algorithmic_flops_per_launch": fl_rec * B * T, "launch_ms": rec_ms,
                "share_of_step": rec_ms / (ms_dev / args.steps) if ms_dev > 0 else None,
                "stage_ms": {"mask_pad": float(stage_ms[0]), "projection_gemm_before_recurrence": float(stage_ms[1]),
                             "recurrence": rec_ms, "recon_mask_gemm": float(stage_ms[3])},
                "stage_note": "pipelined projection: only the projection of the first T/6 frames precedes the persistent "
                              "kernel; the rest of that GEMM runs on the SMs the recurrence leaves free and is inside "
                              "the recurrence interval (DRNMF_FWD_OVERLAP=0 = serial order, same bits)"}

    config = {"workload": "configs[1]: DR-NMF forward via enhance.py path, %d layers, R=%d, %d bins, mask + iSTFT, "
                          "batch of %d synthetic 3 s utterances per GPU" % (K, R, F, B),
              "B_per_gpu": B, "T": T, "F": F, "R": R, "K_layers": K, "N_fft": N, "hop": hop,
              "recurrence": rec_cfg,
              "stft_analysis": "outside both timed regions (features are computed once, as in the reference's cached flow)",
              "l2": "inputs + per-step intermediates (%.2f GB of projections, %.0f MB of S_k hi/lo) exceed the 126 MB L2; "
                    "no explicit flush" % (B * T * K * eng.Rp * 4 / 1e9, 2 * (K - 1) * eng.Rp * eng.Rp * 4 / 1e6)}

    pipe_note = ("3xTF32 issues three tensor passes per useful product: pipe_frac_3xtf32 = 3 x useful / peak.  At B=64 the "
                 "T x K chain of %d dependent all-to-all steps is latency-bound (%.1f us per step against a %.2f us MMA "
                 "floor); the throughput batches below are where the tensor bound applies" %
                 (T * (K - 1), 1e3 * rec_ms / max(T * (K - 1), 1), 1e6 * fl_rec * B * 3 / (tf32_peak * 1e12) / max(K - 1, 1) / 1.0))
    roofline["note"] = pipe_note

    if not args.no_throughput and rank == 0 and world == 1:
        thr = {}
        for Bt, Tt in ((512, T), (2048, min(T, 64))):
            reps = (Bt + B - 1) // B
            xt = x_dev[:, :Tt].repeat(reps, 1, 1)[:Bt].contiguous()
            irm_t = torch.empty((Bt, Tt, F), dtype=torch.float32, device=dev)
            fn = lambda: eng.forward(xt, want_H=False, irm_out=irm_t)
            n = 3
            ms_t, _, st_t = timed(fn, n, 1, with_stage=True)
            useful = fl_rec * Bt * Tt / (float(st_t[2]) / 1e3) / 1e12
            thr["B%d" % Bt] = {"B": Bt, "T": Tt, "frames_per_s_forward_only": Bt * Tt * n / (ms_t / 1e3),
                               "recurrence_ms": float(st_t[2]), "recurrence_useful_tflops": useful,
                               "pipe_frac_3xtf32": 3.0 * useful / tf32_peak, "frac": useful / tf32_peak,
                               "recurrence": eng.recurrent_config()}
            del xt, irm_t
            eng._ws = None
            torch.cuda.empty_cache()
        config["throughput_mode"] = thr

    parity = None
    if not args.no_parity and rank == 0 and world == 1:
        # one untimed step against the float64 oracle on the first utterances of the batch (utterances are independent)
        import oracle as O
        n_par = min(8, B)
        Hd, irm_d = eng.forward(x_dev[:n_par].contiguous())
        pp = dict(p)
        e7 = np.float32(1e-7)
        pp["log_U1"] = np.log(e7 + np.eye(R, dtype=np.float32)); pp["log_Uk"] = np.log(e7 + np.zeros((R, R), dtype=np.float32))
        xo = x_dev[:n_par].cpu().numpy()
        Ho, irmo = O.drnmf_forward(xo, pp, dtype=np.float64)
        full_H, full_irm = eng.forward(x_dev)          # the bench plan itself (B utterances): same rows must match too
        rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
        win = O.sqrt_hann(N)
        audio_full = plan.run(x_dev, stack).cpu().numpy()
        sdr_d = []
        st_np = stack.cpu().numpy()
        for b in range(n_par):
            ref = O.reconstruct_x(st_np[:, b * T:(b + 1) * T].astype(np.float64), hop, win.astype(np.float64), mask=irmo[b].T, dtype=np.float64)[0]
            clean = pairs[b][1]
            m = min(len(clean), ref.size)
            sdr_d.append(abs(O.sdr_db(audio_full[b, :m], clean[:m]) - O.sdr_db(ref[:m], clean[:m])))
        parity = {"against": "oracle/ (numpy float64 restatement of the reference), %d utterances x %d frames" % (n_par, T),
                  "H_rel_fro": rel(full_H[:n_par].cpu().numpy(), Ho), "irm_rel_fro": rel(full_irm[:n_par].cpu().numpy(), irmo),
                  "H_rel_fro_small_batch_plan": rel(Hd.cpu().numpy(), Ho), "sdr_db_max_abs_diff": float(max(sdr_d)),
                  "tolerance": {"H": 1e-4, "irm": 1e-4, "sdr_db": 0.01}}
        parity["ok"] = bool(parity["H_rel_fro"] < 1e-4 and parity["irm_rel_fro"] < 1e-4 and parity["sdr_db_max_abs_diff"] < 0.01)

    if not args.no_extras and rank == 0 and world == 1:
        from drnmf_b200 import training
        ex = {}
        Bt = args.train_batch
        xt, yt = x_dev[:Bt].contiguous(), (x_dev[:Bt] * 0.5).contiguous()
        eng_t = engine.DrnmfEngine(F, R, K)
        tr = training.DeviceTrainer(eng_t, p, learning_rate=1e-4)
        tr.train_on_batch(xt, yt)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            tr.train_on_batch(xt, yt)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        ex["training_step"] = {"B": Bt, "T": T, "ms": 1e3 * dt, "frames_per_s": Bt * T / dt,
                               "forward": eng_t.recurrent_config(), "backward": eng_t.recurrent_config(backward=True),
                               "note": "forward with stored activations + loss + BPTT + weight-gradient GEMMs + parameter chain "
                                       "+ fused Adam + rebuild of the derived weights (see --workload train for the N-GPU line)"}
        del tr, eng_t
        torch.cuda.empty_cache()
        for n_mu, iters in ((22528, 10), (args.mu_frames, 20)):
            try:
                g = torch.Generator(device=dev).manual_seed(3)
                V = torch.rand(F, n_mu, device=dev, generator=g) * 4
                Wm = torch.rand(F, R, device=dev, generator=g) + 0.1
                Hm = torch.rand(R, n_mu, device=dev, generator=g) + 0.1
                engine.snmf_mu_ed(V, Wm, Hm, 1.0, 2)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                engine.snmf_mu_ed(V, Wm, Hm, 1.0, iters)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / iters
                useful = 12.0 * F * R * n_mu / dt / 1e12
                ex["snmf_mu_ed_n%d" % n_mu] = {"F": F, "R": R, "n_frames": n_mu, "ms_per_iteration": 1e3 * dt, "useful_tflops": useful,
                                               "frac": useful / tf32_peak, "pipe_frac_3xtf32": 3.0 * useful / tf32_peak,
                                               "ms_per_100_iterations": 1e5 * dt,
                                               "note": "W and H updated, explicit inits; one host sync per iteration (convergence test)"}
                del V, Wm, Hm
                torch.cuda.empty_cache()
            except Exception as e:      # the extras never take the headline line down
                ex["snmf_mu_ed_n%d" % n_mu] = {"error": repr(e)[:200]}
        # STFT analysis / masked synthesis of the 64-utterance batch against the HBM roofline
        try:
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            for _ in range(2):
                engine.stft_mag(audio, list(offs), lens, N, hop)
            evs[0].record()
            for _ in range(10):
                engine.stft_mag(audio, list(offs), lens, N, hop)
            evs[1].record()
            for _ in range(10):
                _lib_check_istft(engine, plan, stack)
            evs[2].record()
            torch.cuda.synchronize()
            an_bytes = audio.numel() * 4 + stack.numel() * 4 + mag.numel() * 4
            sy_bytes = stack.numel() * 4 + plan.irm.numel() * 4 + plan.audio.numel() * 4
            t_an, t_sy = evs[0].elapsed_time(evs[1]) / 10, evs[1].elapsed_time(evs[2]) / 10
            ex["stft"] = {"analysis_ms": t_an, "analysis_GBps": an_bytes / t_an / 1e6, "analysis_frac_hbm": an_bytes / t_an / 1e6 / pk["hbm_gbs"],
                          "mask_istft_ms": t_sy, "mask_istft_GBps": sy_bytes / t_sy / 1e6, "mask_istft_frac_hbm": sy_bytes / t_sy / 1e6 / pk["hbm_gbs"],
                          "note": "includes the host-side launch + table setup of engine.stft_mag; algorithmic bytes = audio + [Re;Im] "
                                  "stack + magnitudes (analysis), stack + mask + audio (synthesis)"}
        except Exception as e:
            ex["stft"] = {"error": repr(e)[:200]}
        config["extras"] = ex

    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if tj.get("B") == B and tj.get("T") == T and tj.get("R") == R and tj.get("K_layers") == K:
            traffic = tj["dram_bytes_per_launch"]
            roofline["traffic_source"] = tj.get("source")
    except Exception:
        pass
    roofline["traffic"] = traffic

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from threadpoolctl import threadpool_info
            cores = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
        except Exception:
            cores = os.cpu_count() or 1
        n_cpu = min(args.cpu_utts, B)
        fps, dt, _ = cpu_forward_frames_per_s(pairs, p, args, n_cpu)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "%d of the %d utterances of the batch, one pass (%.1f s): numpy restatement of the reference "
                         "(oracle/), float32, BLAS threads" % (n_cpu, B, dt)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (tcgen05 kind::tf32 x3 error-compensated products, fp32 accumulate)",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(x_host.numel() * 4 + stack_host.numel() * 4 + frames_host.numel() * 4),
                    "d2h_bytes_per_step": int(out_host.numel() * 4)},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
