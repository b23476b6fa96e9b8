"""ctypes binding of include/drnmf.h.  No compute happens in Python and there is no fallback:
if libdrnmf.so is missing, or a compute call finds no sm_100 device, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DRNMF_LIB") or os.path.join(_HERE, "libdrnmf.so")      # DRNMF_LIB: debugging aid (A/B of two builds)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "drnmf.h")

IMPL_TCGEN05 = 0
IMPL_SIMT = 1
FLAG_SQUARE_IRM = 16


class DrnmfError(RuntimeError):
    """A libdrnmf call returned a nonzero status (the reference constructs and drops such errors, snmf.py:105)."""

    def __init__(self, code, msg):
        super().__init__("libdrnmf error %d: %s" % (code, msg))
        self.code = code


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p)
LAYER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)

_lib = None


def header_symbols():
    """Every function name include/drnmf.h declares."""
    with open(HEADER_PATH) as f:
        src = f.read()
    return sorted(set(re.findall(r"DRNMF_API[^;(]*?\b(drnmf_\w+)\s*\(", src)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libdrnmf.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "dr-nmf_b200/csrc/build.sh; there is no Python/CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
    sig = {
        "drnmf_version": (i32, []),
        "drnmf_last_error": (C.c_char_p, []),
        "drnmf_launch_count": (C.c_ulonglong, []),
        "drnmf_create": (i32, [C.POINTER(vp), i32, i32, i32, i32]),
        "drnmf_destroy": (i32, [vp]),
        "drnmf_set_params": (i32, [vp, vp, i32, vp, i32, i32, vp, i32, vp, vp, vp, f32, f32, f32, f32, vp]),
        "drnmf_workspace_bytes": (sz, [vp, i32, i32]),
        "drnmf_forward": (i32, [vp, vp, i32, i32, f32, vp, vp, vp, sz, vp]),
        "drnmf_stage_times": (i32, [vp, C.POINTER(f32)]),
        "drnmf_recurrent_config": (i32, [vp, C.POINTER(i32)]),
        "drnmf_recurrent_config2": (i32, [vp, i32, C.POINTER(i32)]),
        "drnmf_debug_inject_error": (i32, [vp, i32, vp]),
        "drnmf_get_derived": (i32, [vp, i32, i32, vp, vp]),
        "drnmf_padded_dims": (i32, [vp, C.POINTER(i32), C.POINTER(i32)]),
        "drnmf_stft_frames": (i32, [i32, i32, i32]),
        "drnmf_stft_mag": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i64, vp, vp, vp]),
        "drnmf_mask_istft": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i64, vp, vp, sz, vp]),
        "drnmf_istft_workspace_bytes": (sz, [i64, i32]),
        "drnmf_enhance_host": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, f32, vp, vp, sz, vp]),
        "drnmf_enhance_workspace_bytes": (sz, [vp, i32, i32, i32, i32]),
        "drnmf_snmf_mu_ed": (i32, [i32, i32, i32, vp, vp, vp, vp, vp, f32, i32, f32, vp, vp, C.POINTER(i32), i32, vp, sz, vp]),
        "drnmf_snmf_workspace_bytes": (sz, [i32, i32, i32]),
        "drnmf_snmf_mu_ed_dist": (i32, [i32, i32, i32, vp, vp, vp, vp, vp, f32, i32, f32, vp, vp, C.POINTER(i32), i32, vp, sz, vp,
                                        ALLREDUCE_FN, vp]),
        "drnmf_snmf_mu_beta": (i32, [i32, i32, i32, f32, vp, vp, vp, vp, vp, f32, i32, f32, vp, vp, C.POINTER(i32), i32, vp, sz,
                                     vp, ALLREDUCE_FN, vp]),
        "drnmf_snmf_beta_workspace_bytes": (sz, [i32, i32, i32, f32]),
        "drnmf_snmf_irm": (i32, [i32, i32, i32, i32, vp, vp, vp, i32, vp, sz, vp]),
        "drnmf_snmf_irm_workspace_bytes": (sz, [i32, i32, i32]),
        "drnmf_ista_ed": (i32, [i32, i32, i32, vp, vp, vp, f32, f32, i32, i32, vp, sz, vp]),
        "drnmf_ista_workspace_bytes": (sz, [i32, i32, i32]),
        "drnmf_loss_and_grads": (i32, [vp, vp, vp, i32, i32, f32, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
        "drnmf_train_workspace_bytes": (sz, [vp, i32, i32]),
        "drnmf_loss_and_grads_cb": (i32, [vp, vp, vp, i32, i32, f32, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp, LAYER_FN, vp]),
        "drnmf_set_training_loss": (i32, [vp, i32, f32]),
        "drnmf_forward_all_hidden": (i32, [vp, vp, i32, i32, f32, vp, vp, sz, vp]),
        "drnmf_adam_step": (i32, [vp, vp, vp, vp, vp, sz, f32, f32, f32, f32, f32, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise DrnmfError(rc, load().drnmf_last_error().decode("utf-8", "replace"))
