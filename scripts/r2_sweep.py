"""Round-2 operating-point sweep of the persistent recurrence: batch tile NB x batch groups G x publish mode, at the
bench batch (B=64), the training batch (B=32) and throughput batches.  The plan is chosen per launch from the
DRNMF_REC_* environment variables, so one process can walk the whole table.  Every configuration is compared with the
first one of its batch size (fixed reduction order -> bitwise equality expected)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drnmf_b200 import engine, synth

F, R, K = 513, int(os.environ.get("SWEEP_R", "1000")), int(os.environ.get("SWEEP_K", "25"))
T = int(os.environ.get("SWEEP_T", "193"))
p = synth.model_params(F, R, K)
p["log_U1"], p["log_Uk"] = synth.structured_u_init()
eng = engine.DrnmfEngine(F, R, K)
eng.set_params(p)
fl_rec = 2.0 * R * R * (K - 1)


def setenv(**kw):
    for k in ("DRNMF_REC_NB", "DRNMF_REC_G", "DRNMF_REC_PUB", "DRNMF_REC_DEBUG", "DRNMF_REC_KS", "DRNMF_REC_VERBOSE", "DRNMF_REC_TRACE",
              "DRNMF_REC_HST", "DRNMF_REC_RST", "DRNMF_REC_WST", "DRNMF_REC_H2D", "DRNMF_REC_LLT", "DRNMF_REC_LL", "DRNMF_REC_NOSYM", "DRNMF_REC_NOSPLIT"):
        os.environ.pop(k, None)
    for k, v in kw.items():
        if v is not None:
            os.environ["DRNMF_REC_" + k] = str(v)


def run(B, Tn, reps=3, **kw):
    g = torch.Generator(device="cuda").manual_seed(B)
    x = torch.rand(B, Tn, F, device="cuda", generator=g) * 4
    setenv(**kw)
    best = None
    H = None
    try:
        for _ in range(reps):
            H, _ = eng.forward(x, want_irm=False)
            torch.cuda.synchronize()
            ms = eng.stage_times()[2]
            best = ms if best is None else min(best, ms)
    except Exception as e:
        print("B=%d %s FAILED: %s" % (B, kw, e), flush=True)
        return None
    cfg = eng.recurrent_config()
    steps = Tn * (K - 1)
    tf = fl_rec * B * Tn / (best / 1e3) / 1e12
    print("B=%4d T=%3d req=%-32s plan=KS%d NB%d G%d tiles%d W%d H%d R%d | rec %8.3f ms  %6.2f us/step  %7.1f kframes/s  %6.1f TF/s useful" % (
        B, Tn, kw, cfg["KS"], cfg["NB"], cfg["groups"], cfg["n_tiles"], cfg["WST"], cfg["HST"], cfg["RST"], best, 1e3 * best / steps,
        B * Tn / best, tf), flush=True)
    return H


def sweep(B, Tn, configs):
    ref = None
    for kw in configs:
        H = run(B, Tn, **kw)
        if H is None:
            continue
        if ref is None:
            ref = H.clone()
        else:
            same = torch.equal(H, ref)
            err = float((H - ref).abs().max() / ref.abs().max())
            print("      vs first config: bitwise %s, max rel %.2e" % (same, err), flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "b64"):
    sweep(64, T, [dict(VERBOSE=1), dict(KS=4, NB=16, G=4), dict(KS=4, NB=32, G=2), dict(KS=4, NB=64, G=1), dict(KS=4, NB=16, G=2),
                  dict(KS=8, NB=64, G=1), dict(KS=8, NB=32, G=1), dict(KS=2, NB=16, G=4), dict(KS=4, NB=16, G=4, PUB="thread")])
if which in ("all", "b32"):
    sweep(32, T, [dict(), dict(KS=4, NB=16, G=2), dict(KS=4, NB=32, G=1), dict(KS=8, NB=32, G=1)])
if which in ("all", "thr"):
    sweep(512, 48, [dict(VERBOSE=1), dict(KS=4, NB=64, G=4), dict(KS=4, NB=32, G=4), dict(KS=4, NB=64, G=4, PUB="direct"),
                    dict(KS=4, NB=32, G=4, PUB="direct"), dict(KS=8, NB=64, G=1), dict(KS=2, NB=32, G=8)])
    sweep(2048, 12, [dict(), dict(KS=4, NB=32, G=4), dict(KS=8, NB=64, G=1)])
    sweep(128, 96, [dict(), dict(KS=4, NB=32, G=4), dict(KS=4, NB=64, G=2), dict(KS=8, NB=64, G=1)])
    sweep(256, 96, [dict(), dict(KS=4, NB=32, G=4)])
if which in ("all", "dbg"):
    print("---- per-role cycle counters (stderr) ----", flush=True)
    for B, Tn in ((64, 40), (32, 40), (512, 12), (2048, 6)):
        sys.stderr.write("\n## B=%d (default plan)\n" % B); sys.stderr.flush()
        run(B, Tn, reps=1, DEBUG=1)
if which in ("final",):
    sweep(64, T, [dict(), dict(KS=4, NB=16, G=4), dict(NOSYM=1)])
    sweep(32, T, [dict(), dict(LL=0)])
    sweep(16, T, [dict(), dict(LL=0)])
    sweep(48, T, [dict()])
    sweep(128, 96, [dict(), dict(KS=8, NB=64, G=1)])
    sweep(256, 96, [dict()])
    sweep(512, 48, [dict(), dict(PUB="thread")])
    sweep(768, 32, [dict()])
    sweep(1024, 24, [dict(), dict(NOSPLIT=1)])
    sweep(2048, 12, [dict(), dict(NOSPLIT=1), dict(KS=4, NB=64, PUB="thread")])
    sweep(4096, 6, [dict()])
if which in ("b64x",):
    sweep(64, T, [dict(), dict(HST=2, RST=1, WST=2), dict(HST=2, RST=2, WST=2), dict(HST=4, RST=1, WST=2), dict(HST=2, RST=1, WST=4),
                  dict(HST=2, RST=1, WST=3), dict(HST=3, RST=1, WST=2)])
    sweep(512, 48, [dict(), dict(NOSYM=1), dict(NOSYM=1, WST=2), dict(NOSYM=1, HST=2, WST=4)])
    sweep(2048, 12, [dict(), dict(NOSYM=1)])
if which in ("llt",):
    sweep(64, T, [dict(), dict(LLT=2), dict(LLT=2, WST=4), dict(LLT=2, HST=2), dict(KS=4, NB=32, G=2), dict(KS=4, NB=32, G=2, WST=2),
                  dict(KS=4, NB=16, G=2, LLT=2), dict(KS=8, NB=16, G=1, LLT=4), dict(KS=8, NB=16, G=1)])
    sweep(48, T, [dict(), dict(LLT=2)])
if which in ("thr2",):
    sweep(512, 48, [dict(PUSH=0, DEFER=0), dict(), dict(PUSH=1, DEFER=0), dict(PUSH=0, DEFER=1), dict(RST=3, HST=2), dict(RST=2, HST=2, WST=4),
                    dict(KS=4, NB=32, G=4), dict(KS=8, NB=64, G=1)])
    sweep(2048, 12, [dict(PUSH=0, DEFER=0), dict(), dict(PUSH=1, DEFER=0), dict(PUSH=0, DEFER=1), dict(RST=3, HST=2), dict(RST=2, HST=2, WST=4)])
    sweep(256, 96, [dict(PUSH=0, DEFER=0), dict(), dict(KS=4, NB=32, G=4), dict(KS=4, NB=32, G=4, PUSH=1, DEFER=1)])
    sweep(64, T, [dict(), dict(DEFER=1), dict(PUSH=1), dict(PUSH=1, DEFER=1)])
if which in ("b32dbg",):
    for kw, Tn in ((dict(), 193), (dict(), 8), (dict(), 2), (dict(LL=0), 193), (dict(NOSYM=1), 193), (dict(WST=2), 193), (dict(DEBUG=1), 40), (dict(NB=16), 193)):
        sys.stderr.write("\n## B=32 T=%d %s\n" % (Tn, kw)); sys.stderr.flush()
        run(32, Tn, reps=1, **kw)
if which in ("ab",):
    for B, Tn in ((64, T), (32, T), (16, T), (512, 48), (768, 32), (1024, 24), (2048, 12), (4096, 6)):
        run(B, Tn)
if which in ("ks2",):
    sweep(2048, 12, [dict(), dict(KS=2, NB=64, VERBOSE=1), dict(KS=2, NB=64, G=8), dict(KS=2, NB=64, G=4), dict(KS=2, NB=32)])
    sweep(512, 48, [dict(), dict(KS=2, NB=64), dict(KS=2, NB=64, G=8), dict(KS=2, NB=64, G=4), dict(KS=2, NB=32)])
    sweep(1024, 24, [dict(), dict(KS=2, NB=64)])
    sys.stderr.write("\n## B=2048 KS=2\n"); sys.stderr.flush()
    run(2048, 6, reps=1, DEBUG=1, KS=2, NB=64)
if which in ("rings",):
    sweep(2048, 12, [dict(), dict(HST=2, WST=4), dict(HST=2, WST=3), dict(HST=3, WST=2, RST=1), dict(HST=2, WST=2)])
    sweep(1024, 24, [dict(), dict(HST=2, WST=4)])
if which in ("rings2",):
    sweep(512, 48, [dict(), dict(HST=2, WST=2), dict(HST=2, WST=3), dict(HST=3, WST=2), dict(HST=2, WST=4)])
    sweep(256, 96, [dict(), dict(HST=2, WST=2), dict(HST=2, WST=4)])
    sweep(128, 96, [dict(), dict(HST=2, WST=2), dict(HST=4, WST=2), dict(HST=2, WST=4)])
    sweep(4096, 6, [dict(), dict(HST=2, WST=2)])
if which in ("dbg2",):
    for B, Tn in ((2048, 6), (512, 12)):
        sys.stderr.write("\n## B=%d (default plan)\n" % B); sys.stderr.flush()
        run(B, Tn, reps=1, DEBUG=1)
if which in ("re1",):     # re-entry: the latency-regime alternatives once more (their earlier logs were lost with the container)
    sweep(64, T, [dict(), dict(KS=4, NB=32, G=2, VERBOSE=1), dict(KS=4, NB=32, G=2, WST=2), dict(KS=4, NB=32, G=2, LL=0), dict(KS=4, NB=16, G=4)])
    sweep(32, T, [dict(), dict(KS=4, NB=32, G=1), dict(KS=4, NB=16, G=2)])
if which in ("re2",):
    sweep(64, T, [dict(), dict(KS=8, NB=16, G=1), dict(KS=8, NB=16, G=1, HST=4, RST=2, WST=2), dict(KS=8, NB=16, G=1, HST=8, RST=4, WST=2),
                  dict(KS=8, NB=16, G=1, HST=4, RST=4, WST=4), dict(KS=8, NB=16, G=1, NOSYM=1), dict(KS=8, NB=16, G=1, LLT=4)])
    sweep(48, T, [dict(), dict(KS=8, NB=16, G=1)])
if which in ("re3",):     # event traces of two consecutive steps (4 items) at B = 64: flags vs self-validating exchange with two tiles
    for kw in (dict(), dict(LLT=2), dict(PUB="thread")):
        sys.stderr.write("\n## B=64 %s\n" % kw); sys.stderr.flush()
        run(64, 40, reps=1, DEBUG=1, TRACE="400:404", **kw)
    sys.stderr.write("\n## B=32\n"); sys.stderr.flush()
    run(32, 40, reps=1, DEBUG=1, TRACE="200:202")
if which in ("re4",):
    sys.stderr.write("\n## B=64\n"); sys.stderr.flush()
    run(64, 40, reps=1, DEBUG=1, TRACE="400:402")
if which in ("crash",):
    run(int(os.environ.get("CRASH_B", "32")), int(os.environ.get("CRASH_T", "193")), reps=2)
if which in ("trace2",):
    for B, Tn, nt, kw in ((64, 40, 1, dict(KS=8, NB=64, G=1)), (64, 40, 1, dict())):
        sys.stderr.write("\n## B=%d %s\n" % (B, kw)); sys.stderr.flush()
        run(B, Tn, reps=1, DEBUG=1, TRACE="%d:%d" % (200 * nt, 200 * nt + 2 * nt), **kw)
if which in ("trace",):
    for B, Tn, nt, kw in ((64, 40, 1, dict()), (64, 40, 2, dict(KS=8, NB=32, G=1)), (64, 40, 2, dict(KS=8, NB=32, G=1, PUB="direct")),
                          (64, 40, 1, dict(KS=8, NB=64, G=1)), (512, 12, 2, dict(PUB="direct")), (2048, 6, 8, dict(PUB="direct"))):
        sys.stderr.write("\n## B=%d %s\n" % (B, kw)); sys.stderr.flush()
        run(B, Tn, reps=1, DEBUG=1, TRACE="%d:%d" % (200 * nt, 200 * nt + 2 * nt), **kw)
setenv()
