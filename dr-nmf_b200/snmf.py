"""Host-side mirror of the reference's snmf.py: sparse_nmf_matlab (the chunk driver, snmf.py:9-85) and
sparse_nmf_matlab_on_chunk (snmf.py:88-113), with the MATLAB subprocess + .mat file IPC replaced by the in-process
CUDA solver behind drnmf_snmf_mu_beta (sparseNMF/sparse_nmf_gpu.m: cf = 'ed' - every shipped config, enhance.py:568,590 -
'kl', 'is' or a numeric beta).

Differences that are deliberate: errors are
raised instead of constructed-and-dropped (snmf.py:105-106); missing initialisers are drawn from numpy's
default_rng(random_seed) because MATLAB's legacy rand('seed') stream (sparse_nmf_gpu.m:119) cannot be reproduced;
an array init_h is sliced per chunk (the reference would hand MATLAB a mis-sized matrix)."""
from __future__ import annotations

import copy
import hashlib
import json
import os

import numpy as np
import torch

from . import engine as _engine


def sparse_nmf_matlab(V, params, verbose=True, useGPU=True, gpuIndex=1, save_H=True):
    """snmf.py:9-85.  V (n_feats, n_frames) nonnegative; returns (W, H or None, obj) with obj = {'cost','div'} (plus
    'obj_snmf_per_chunk' when chunked)."""
    params_copy = copy.deepcopy(params)
    n_feats, n_frames = V.shape
    r = int(params["r"])
    frame_batch_size = int(float(700000) * (200.0 / float(r)))          # snmf.py:33-35
    n_chunks = int(np.ceil(float(n_frames) / float(frame_batch_size)))
    H = np.zeros((r, n_frames)) if save_H else None
    obj_snmf = {"obj_snmf_per_chunk": []}
    initial_cost = final_cost = initial_div = final_div = 0.0
    init_h_full = params_copy.get("init_h", None)
    W = None
    for i in range(n_chunks):
        if verbose:
            print("sparse NMF: processing chunk %d of %d..." % (i + 1, n_chunks))
        s, e = i * frame_batch_size, (i + 1) * frame_batch_size
        pc = dict(params_copy)
        if isinstance(init_h_full, np.ndarray):
            pc["init_h"] = init_h_full[:, s:e]
        W, H_tmp, obj_tmp = sparse_nmf_matlab_on_chunk(V[:, s:e], pc, verbose=verbose, gpuIndex=gpuIndex)
        if "w_update_ind" in params_copy:                                # snmf.py:60-64
            idx = np.where(np.asarray(params_copy["w_update_ind"]).astype(bool))[0]
            params_copy["init_w"] = np.array(params_copy["init_w"], dtype=W.dtype)
            params_copy["init_w"][:, idx] = W[:, idx]
        else:
            params_copy["init_w"] = W
        obj_snmf["obj_snmf_per_chunk"].append(obj_tmp)
        initial_cost += obj_tmp["cost"][0]; initial_div += obj_tmp["div"][0]
        final_cost += obj_tmp["cost"][-1]; final_div += obj_tmp["div"][-1]
        if save_H:
            H[:, s:e] = H_tmp
    obj_snmf["cost"] = [initial_cost, final_cost]
    obj_snmf["div"] = [initial_div, final_div]
    if n_chunks == 1:
        obj_snmf = obj_snmf["obj_snmf_per_chunk"][0]
    return W, H, obj_snmf


def sparse_nmf_matlab_on_chunk(V, params, verbose=True, useGPU=True, gpuIndex=1, impl=None):
    """snmf.py:88-113 without MATLAB: parameter handling of sparse_nmf_gpu.m:72-161, solver on the GPU."""
    if not useGPU:
        raise NotImplementedError("there is no CPU solver: the sparse-NMF path runs on the B200 only")
    cf = params.get("cf", "kl")                                               # sparse_nmf_gpu.m:100-115
    beta = {"is": 0.0, "kl": 1.0, "ed": 2.0}.get(cf, float(params.get("beta", 1.0)))
    V = np.asarray(V)
    m, n = V.shape
    rng = np.random.default_rng(int(params.get("random_seed", 1)))
    if "init_w" not in params or params["init_w"] is None:
        r = int(params["r"])
        w = rng.random((m, r))
    else:
        w = np.array(params["init_w"], dtype=np.float64)
        if "r" in params and w.shape[1] < int(params["r"]):               # sparse_nmf_gpu.m:129-133
            w = np.concatenate([w, rng.random((m, int(params["r"]) - w.shape[1]))], axis=1)
        r = w.shape[1]
    ih = params.get("init_h", None)
    if ih is None:
        h = rng.random((r, n))
    elif isinstance(ih, str) and ih == "ones":
        h = np.ones((r, n))
    else:
        h = np.array(ih, dtype=np.float64)
    sp = np.asarray(params.get("sparsity", 0.0), dtype=np.float64)
    if sp.size != 1:
        raise NotImplementedError("per-entry sparsity matrices (sparse_nmf_gpu.m:157-161) are not used by the path")
    dev = torch.device("cuda", torch.cuda.current_device())
    Vd = torch.as_tensor(np.ascontiguousarray(V, dtype=np.float32), device=dev)
    Wd = torch.as_tensor(np.ascontiguousarray(w, dtype=np.float32), device=dev)
    Hd = torch.as_tensor(np.ascontiguousarray(h, dtype=np.float32), device=dev)
    cost, div = _engine.snmf_mu_ed(Vd, Wd, Hd, float(sp.reshape(())), int(params.get("max_iter", 100)),
                                   float(params.get("conv_eps", 0.0)), params.get("w_update_ind", None),
                                   params.get("h_update_ind", None), impl=impl, beta=beta)
    W = Wd.cpu().numpy().astype(V.dtype)
    H = Hd.cpu().numpy().astype(V.dtype)
    return W, H, {"cost": cost, "div": div}


class _NumpyEncoder(json.JSONEncoder):
    """enhance.py:60-71 (MyEncoder)."""
    def default(self, obj):
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        return super().default(obj)


def get_snmf_savefile(params_snmf, path_dicts=""):
    """enhance.py:74-79: path_dicts + 'W_noisy_<md5 of the sorted-key JSON of the parameters>_sparsity%.3f', with '.npz'
    in place of the reference's '.hkl' (hickle is not available; same stem, so caches are found by the same key)."""
    h = hashlib.md5(json.dumps(params_snmf, sort_keys=True, cls=_NumpyEncoder).encode()).hexdigest()
    return path_dicts + "W_noisy_" + h + ("_sparsity%.3f.npz" % params_snmf["sparsity"])


def _load_or_none(path, flag_recompute):
    if path is None or flag_recompute or not os.path.exists(path):
        return None
    z = np.load(path, allow_pickle=True)
    H = z["H"] if "H" in z.files and z["H"].ndim == 2 else None
    return z["W"], H, {"cost": z["cost"], "div": z["div"]}


def _dump(path, W, H, obj, save_H):
    if path is None:
        return
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    np.savez(path, W=W, H=(H if (save_H and H is not None) else np.zeros(0)), cost=np.asarray(obj["cost"]),
             div=np.asarray(obj["div"]))


def train_snmf(clean_frames, noisy_frames, params_snmf, noise_init=None, verbose=False, save_H=True, flag_recompute=False,
               path_dicts=None):
    """enhance.py:81-135: stage 1 learns r clean atoms, stage 2 learns [W_clean, W_noise] on the noisy frames with the
    clean atoms frozen (w_update_ind).  noise_init replaces np.random.rand(*W.shape).  With path_dicts the two
    dictionaries are cached under the reference's names (get_snmf_savefile; 'noisy' -> 'clean' for stage 1) and
    reloaded unless flag_recompute."""
    f_noisy = get_snmf_savefile(params_snmf, path_dicts) if path_dicts is not None else None
    f_clean = f_noisy.replace("noisy", "clean") if f_noisy else None
    got = _load_or_none(f_clean, flag_recompute)
    if got is None:
        W, H, obj = sparse_nmf_matlab(clean_frames, params_snmf, verbose=verbose, save_H=save_H)
        _dump(f_clean, W, H, obj, save_H)
    else:
        W, H, obj = got
    r = int(params_snmf["r"])
    if noise_init is None:
        noise_init = np.random.default_rng(7654).random(W.shape)
    W_init = np.concatenate((W, np.asarray(noise_init, dtype=np.float32)), axis=1)
    idx_update = np.concatenate((np.zeros(r, dtype=bool), np.ones(r, dtype=bool)))
    p2 = copy.deepcopy(params_snmf)
    p2.update({"r": 2 * r, "init_w": W_init, "w_update_ind": idx_update})
    got = _load_or_none(f_noisy, flag_recompute)
    if got is None:
        W_noisy, H_noisy, obj_noisy = sparse_nmf_matlab(noisy_frames, p2, verbose=verbose, save_H=save_H)
        _dump(f_noisy, W_noisy, H_noisy, obj_noisy, save_H)
    else:
        W_noisy, H_noisy, obj_noisy = got
    obj_noisy["cost"] = np.squeeze(obj_noisy["cost"])
    obj_noisy["div"] = np.squeeze(obj_noisy["div"])
    return W_noisy, H_noisy, obj_noisy


def snmf_infer(x_frames, W_noisy, params_snmf, max_iter=200, verbose=False):
    """enhance.py:836-845: activations of a FIXED dictionary on new frames (w_update_ind all false, conv_eps 0,
    max_iter 200).  x_frames (F, n); returns H (2r, n) and the objective trace."""
    p = copy.deepcopy(params_snmf)
    R = W_noisy.shape[1]
    p.update({"r": R, "init_w": np.asarray(W_noisy), "w_update_ind": np.zeros(R, dtype=bool), "conv_eps": 0.0,
              "max_iter": float(max_iter)})
    _, H, obj = sparse_nmf_matlab(x_frames, p, verbose=verbose)
    return H, obj


def snmf_irm(W_noisy, H, r):
    """enhance.py:847-852: irm = S^ / (1e-9 + S^ + N^) with S^ = W_clean H_clean, N^ = W_noise H_noise, as ONE dual-operand
    tcgen05 GEMM with the ratio in its epilogue (drnmf_snmf_irm).  W_noisy (F, 2r), H (2r, n) -> (F, n) float32 numpy."""
    from . import engine as _engine
    dev = torch.device("cuda", torch.cuda.current_device())
    Wd = torch.as_tensor(np.ascontiguousarray(W_noisy, dtype=np.float32), device=dev)
    Hn = np.ascontiguousarray(H, dtype=np.float32)
    n = Hn.shape[1]
    pad = (-n) % 4
    if pad:
        Hn = np.concatenate([Hn, np.zeros((Hn.shape[0], pad), np.float32)], axis=1)
    irm = _engine.snmf_irm(Wd, torch.as_tensor(Hn, device=dev), r)
    return irm[:, :n].cpu().numpy()
