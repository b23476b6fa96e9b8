"""Time the STFT analysis and the masked iSTFT + overlap-add kernels at the bench shape (64 x 3 s, N=1024, hop=256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drnmf_b200 import engine
B, sec, N, hop = 64, 3.0, 1024, 256
L = int(sec * 16000)
rng = np.random.default_rng(0)
audio = torch.as_tensor(rng.standard_normal(B * L).astype(np.float32), device="cuda")
offs = [b * L for b in range(B)]; lens = [L] * B
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
stack, mag, fidx = engine.stft_mag(audio, offs, lens, N, hop)
mask = torch.rand_like(mag)
t_a = timed(lambda: engine.stft_mag(audio, offs, lens, N, hop))
t_s = timed(lambda: engine.mask_istft(stack, mask, fidx, N, hop))
F = N // 2 + 1; nf = mag.shape[0]
bytes_a = B * L * 4 + nf * F * 4 * 3          # audio in, stack (2F) + magnitude out
bytes_s = nf * F * 4 * 3 + B * (hop * (nf // B - 1) - N) * 4
print("frames %d  analysis %.1f us (%.2f TB/s algorithmic)  synthesis+OLA %.1f us (%.2f TB/s algorithmic)  [includes host-side table setup per call]"
      % (nf, 1e3 * t_a, bytes_a / t_a / 1e9, 1e3 * t_s, bytes_s / t_s / 1e9))
