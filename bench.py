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
    ap.add_argument("--throughput-batch", type=int, default=0,
                    help="also time a large device-resident batch (reported under config.throughput_mode)")
    ap.add_argument("--extras", action="store_true",
                    help="also time a training step (B=32) and sparse-NMF iterations (reported under config.extras)")
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
def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
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
                "stage_ms": {"mask_pad": float(stage_ms[0]), "projection_gemm": float(stage_ms[1]),
                             "recurrence": rec_ms, "recon_mask_gemm": float(stage_ms[3])}}

    config = {"workload": "configs[1]: DR-NMF forward via enhance.py path, %d layers, R=%d, %d bins, mask + iSTFT, "
                          "batch of %d synthetic 3 s utterances per GPU" % (K, R, F, B),
              "B_per_gpu": B, "T": T, "F": F, "R": R, "K_layers": K, "N_fft": N, "hop": hop,
              "recurrence": rec_cfg,
              "l2": "inputs + per-step intermediates (%.2f GB of projections, %.0f MB of S_k hi/lo) exceed the 126 MB L2; "
                    "no explicit flush" % (B * T * K * eng.Rp * 4 / 1e9, 2 * (K - 1) * eng.Rp * eng.Rp * 4 / 1e6)}

    if args.throughput_batch and rank == 0 and world == 1:
        Bt = args.throughput_batch
        reps = (Bt + B - 1) // B
        xt = x_dev.repeat(reps, 1, 1)[:Bt].contiguous()
        irm_t = torch.empty((Bt, T, F), dtype=torch.float32, device=dev)
        fn = lambda: eng.forward(xt, want_H=False, irm_out=irm_t)
        ms_t, _, st_t = timed(fn, max(2, args.steps // 3), 1, with_stage=True)
        n = max(2, args.steps // 3)
        config["throughput_mode"] = {"B": Bt, "frames_per_s_forward_only": Bt * T * n / (ms_t / 1e3),
                                     "recurrence_ms": float(st_t[2]),
                                     "recurrence_useful_tflops": fl_rec * Bt * T / (float(st_t[2]) / 1e3) / 1e12,
                                     "recurrence": eng.recurrent_config()}

    if args.extras and rank == 0 and world == 1:
        ex = {}
        Bt = 32
        xt, yt = x_dev[:Bt].contiguous(), (x_dev[:Bt] * 0.5).contiguous()
        eng.loss_and_grads(xt, yt)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            eng.loss_and_grads(xt, yt)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        ex["training_step"] = {"B": Bt, "T": T, "ms": 1e3 * dt, "frames_per_s": Bt * T / dt,
                               "note": "forward with stored activations + loss + BPTT + weight-gradient GEMMs + parameter chain"}
        n_mu = 22528
        V = torch.rand(F, n_mu, device=dev) * 4
        Wm = torch.rand(F, R, device=dev) + 0.1
        Hm = torch.rand(R, n_mu, device=dev) + 0.1
        engine.snmf_mu_ed(V, Wm, Hm, 1.0, 2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        engine.snmf_mu_ed(V, Wm, Hm, 1.0, 10)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        ex["snmf_mu_ed"] = {"F": F, "R": R, "n_frames": n_mu, "ms_per_iteration": 1e3 * dt,
                            "useful_tflops": 12.0 * F * R * n_mu / dt / 1e12,
                            "note": "W and H updated; includes one host sync per iteration for the convergence test"}
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
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
