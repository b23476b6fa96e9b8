"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/drnmf.h declares, and
refuses to compute without a B200 (no fallback)."""
import ctypes
import os

import pytest
import torch

from drnmf_b200 import _lib, engine


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _lib.header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libdrnmf.so does not export %s" % n
    assert lib.drnmf_version() >= 100


def test_host_only_entry_points():
    lib = _lib.load()
    # stft_frames follows util.py:184-190 + librosa center=False
    assert lib.drnmf_stft_frames(48000, 512, 128) == 380
    assert lib.drnmf_stft_frames(48000, 1024, 256) == 193
    assert lib.drnmf_stft_frames(48000, 2048, 512) == 99
    assert lib.drnmf_stft_frames(333, 64, 16) == 26
    assert lib.drnmf_istft_workspace_bytes(10, 64) >= 10 * 64 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.drnmf_create(ctypes.byref(h), 33, 16, 3, 0)
    assert rc == 5, "expected DRNMF_ERR_NO_DEVICE"
    assert b"no CPU fallback" in lib.drnmf_last_error()
    with pytest.raises(_lib.DrnmfError):
        engine.DrnmfEngine(33, 16, 3)
    with pytest.raises(TypeError):
        engine.DrnmfEngine.forward(object.__new__(engine.DrnmfEngine), torch.zeros(1, 2, 3))


def test_structured_u_detection():
    import numpy as np
    R = 6
    e = np.float32(1e-7)
    d, o = engine.structured_u(np.log(e + np.eye(R, dtype=np.float32)))
    assert abs(d - 1.0) < 1e-6 and abs(o - 1e-7) < 1e-12
    with pytest.raises(NotImplementedError):
        engine.structured_u(np.random.default_rng(0).standard_normal((R, R)))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's baseline arm) runs without a GPU and prints ONE JSON line with the
    contract keys; it is the only bench leg allowed to execute oracle/."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-utts", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-1500:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_every_export_cites_the_reference_interface_it_replaces():
    """include/drnmf.h: each entry point's comment names the reference file:line it stands in for (the drop-in boundary
    is defined by those citations); only the instrumentation / test hooks have no counterpart in the reference."""
    import re
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(here, "include", "drnmf.h")).read()
    hooks = {"drnmf_version", "drnmf_last_error", "drnmf_launch_count", "drnmf_stage_times", "drnmf_recurrent_config",
             "drnmf_recurrent_config2", "drnmf_debug_inject_error", "drnmf_get_derived", "drnmf_padded_dims"}
    uncited = []
    for m in re.finditer(r"DRNMF_API\s+[\w\s\*]+?\b(drnmf_\w+)\s*\(", text):
        block = text[text[:m.start()].rfind("/*"):m.start()]
        if m.group(1) not in hooks and not re.search(r"\.(py|m):\d+", block):
            uncited.append(m.group(1))
    assert not uncited, uncited
