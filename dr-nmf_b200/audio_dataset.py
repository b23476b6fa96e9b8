"""Host-side mirror of the part of the reference's audio_dataset.py that is on the hot path: the [Re;Im] STFT stack
with its fidx table and reconstruct_x (audio_dataset.py:267-278).  The CHiME2 file handling (taskfiles, HDF5 cache,
MATLAB scoring) is out of scope; AudioDataset here is built from in-memory waveforms."""
from __future__ import annotations

import numpy as np
import torch

from . import engine as _engine
from .util import sqrt_hann


class AudioDataset:
    def __init__(self, x_waveforms, y_waveforms=None, params_stft={"N": 320, "hop": 160, "nch": 1}):
        self.params_stft = dict(params_stft)
        self.params_stft["window"] = sqrt_hann(self.params_stft["N"])            # audio_dataset.py:194
        N, hop = self.params_stft["N"], self.params_stft["hop"]
        dev = torch.device("cuda", torch.cuda.current_device())

        def stacks(waves):
            lens = [len(w) for w in waves]
            offs = np.concatenate([[0], np.cumsum(lens)])[:-1]
            audio = torch.as_tensor(np.concatenate(waves).astype(np.float32), device=dev)
            return _engine.stft_mag(audio, list(offs), lens, N, hop)

        self.x_stack_dev, self.x_mag_dev, self.fidx_dev = stacks(x_waveforms)
        self.x_stack = self.x_stack_dev.cpu().numpy()                             # (2F, total frames), util.py:351
        self.fidx = self.fidx_dev.cpu().numpy().astype(np.int32)                  # (n_files, 2), util.py:335-337
        if y_waveforms is not None:
            self.y_stack_dev, self.y_mag_dev, _ = stacks(y_waveforms)
            self.y_stack = self.y_stack_dev.cpu().numpy()

    def _reconstruct(self, stack_dev, idx, mask):
        N, hop = self.params_stft["N"], self.params_stft["hop"]
        s, e = int(self.fidx[idx, 0]), int(self.fidx[idx, 1])
        sub = stack_dev[:, s:e].contiguous()
        fidx = torch.as_tensor(np.array([[0, e - s]], dtype=np.int64), device=sub.device)
        m = None
        if mask is not None:   # mask (F, frames) as in the reference; the kernel wants (frames, F)
            m = torch.as_tensor(np.ascontiguousarray(np.asarray(mask, np.float32).T), device=sub.device)
        (y,) = _engine.mask_istft(sub, m, fidx, N, hop)
        return y.cpu().numpy()[None, :]

    def reconstruct_x(self, idx, mask=None):
        """audio_dataset.py:267-278: mask tiled over [Re;Im], multiplied, istft_mc(flag_noDiv=1)."""
        return self._reconstruct(self.x_stack_dev, idx, mask)

    def reconstruct_y(self, idx, mask=None):
        return self._reconstruct(self.y_stack_dev, idx, mask)
