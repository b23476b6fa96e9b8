"""torch float64 copy of the oracle's DR-NMF forward + training loss, used only to obtain reference GRADIENTS through
torch.autograd (the reference has no hand-written backward: Theano autodiff through scan, enhance.py:1152).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import numpy as np
import torch

EPS = 1e-7
GRAD_KEYS = ("log_D", "log_alph", "log_lam1", "log_h0", "k_clean", "k_noise")


def forward_loss(x, y, p, mask_value=-1.0, transform_before_irm=None, loss="mse_of_masked", lam1=0.0, return_all_hidden=False):
    """x, y (B,T,F) numpy; p: parameter dict (numpy).  Returns (loss tensor, dict of leaf tensors, H, irm).
    Same arithmetic as oracle.rnn_forward / output_head / training_loss (custom_layers.py:343-375,
    enhance.py:269-305, 1040-1073), structured U."""
    dt = torch.float64
    leaves = {k: torch.tensor(np.asarray(p[k], dtype=np.float64), dtype=dt, requires_grad=True) for k in GRAD_KEYS}
    xt = torch.tensor(np.asarray(x, dtype=np.float64))
    yt = torch.tensor(np.asarray(y, dtype=np.float64))
    B, T, F = xt.shape
    K, _, R = leaves["log_D"].shape
    r = R // 2
    m = (xt != mask_value).any(dim=-1)
    xm = xt * m[..., None]

    def ud(key):
        U = np.exp(np.asarray(p[key], dtype=np.float64))
        return float(U[0, 0]), float(U[0, 1])
    d0, o0 = ud("log_U1")
    dk, ok = ud("log_Uk")
    Wk, Sk, bk = [], [], []
    for k in range(K):
        D = torch.exp(leaves["log_D"][k])
        Dn = D / torch.sqrt(torch.sum(D * D, dim=0, keepdim=True))
        alph = torch.exp(leaves["log_alph"][k])
        lam = torch.exp(leaves["log_lam1"][k])
        Wk.append(Dn / alph)
        bk.append(-torch.ones(R, dtype=dt) * lam / alph)
        Sk.append((torch.eye(R, dtype=dt) - (Dn / alph).T @ Dn).T if k > 0 else None)
    state = torch.nn.functional.softplus(leaves["log_h0"])[None, :].expand(B, R)
    out_prev = torch.zeros(B, R, dtype=dt)
    all_prev = torch.zeros(B, K * R, dtype=dt)
    Hs, Hall = [], []
    for t in range(T):
        prev = state
        psum = prev.sum(dim=1, keepdim=True)
        g = None
        hid = []
        for k in range(K):
            d, o = (d0, o0) if k == 0 else (dk, ok)
            pre = prev * (d - o) + o * psum
            if k > 0:
                pre = pre + g @ Sk[k]
            pre = pre + xm[:, t, :] @ Wk[k]
            g = torch.relu(pre + bk[k])
            hid.append(g)
        mt = m[:, t][:, None]
        out_prev = torch.where(mt, g, out_prev)
        state = torch.where(mt, g, state)
        Hs.append(out_prev)
        if return_all_hidden:      # custom_layers.py:371-374: concatenation of all layers, carried over masked frames
            cat = torch.cat(hid, dim=1)
            all_prev = torch.where(mt, cat, all_prev)
            Hall.append(all_prev)
    H = torch.stack(Hs, dim=1)
    S = H[..., :r] @ torch.exp(leaves["k_clean"])
    N = H[..., r:] @ torch.exp(leaves["k_noise"])
    if transform_before_irm == "square":
        S, N = S * S, N * N
    irm = torch.exp(torch.log(EPS + S) - torch.log(EPS + S + N))
    mf = m.to(dt)
    if loss == "snmf_cost":        # enhance.py:1024-1036: [mse(x_recon, x), mean|h|] with weights [0.5, lam1 * 2r / F]
        per = 0.5 * torch.mean((S + N - xt) ** 2, dim=-1) + lam1 * (R / F) * torch.mean(torch.abs(H), dim=-1)
    else:
        per = torch.mean((xt * irm - yt) ** 2, dim=-1)
    loss = torch.sum(per * mf) / torch.sum(mf)
    if return_all_hidden:
        return loss, leaves, H, irm, torch.stack(Hall, dim=1)
    return loss, leaves, H, irm


def loss_and_grads(x, y, p, **kw):
    loss, leaves, H, irm = forward_loss(x, y, p, **kw)[:4]
    loss.backward()
    return float(loss.detach()), {k: v.grad.numpy() for k, v in leaves.items()}, H.detach().numpy(), irm.detach().numpy()
