"""Host-side mirror of the reference's audio_dataset.py around the hot path: the [Re;Im] STFT stack with its fidx
table, reconstruct_x (audio_dataset.py:267-278), and the data format the network consumes - padded (n_sequences,
maxlen, d) tensors with a mask (reshape_and_pad_stacks, :116-169; get_padded_data_matrix, :369-383; the 'mag' /
'logmag' feature maps and the pad value of load_data, :11-37).  The CHiME2 file handling (taskfiles, HDF5 cache, MATLAB
scoring) is out of scope; AudioDataset here is built from in-memory waveforms."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import engine as _engine
from .util import sqrt_hann


def get_mask_value(config):
    """audio_dataset.py:11-17."""
    if config.get("transform_x") == "mag" or config.get("transform_y") == "logmag":
        return -1.0
    return 0.0


def data_transform(kind):
    """audio_dataset.py:22-37: feature map applied to a [Re;Im] stack (2F, frames) -> (F, frames); identity otherwise."""
    def _mag(x):
        F = x.shape[0] // 2
        return np.sqrt(x[:F] ** 2 + x[F:] ** 2)
    if kind == "mag":
        return _mag
    if kind == "logmag":
        return lambda x: np.log(np.float32(1.0) + _mag(x))
    return lambda x: x


def sequence_table(fidx, maxlen=None):
    """Chunk table behind reshape_and_pad_stacks: rows (start frame, end frame) of every output sequence.  A file longer
    than maxlen is cut into consecutive pieces of maxlen frames; pieces never span two files (audio_dataset.py:120-167)."""
    fidx = np.asarray(fidx, dtype=np.int64)
    maxseq = int((fidx[:, 1] - fidx[:, 0]).max())
    if maxlen is None or maxlen > maxseq:
        maxlen = maxseq
    maxlen = int(maxlen)
    rows = []
    for s, e in fidx:
        if maxlen == maxseq:
            rows.append((s, e))
        else:
            rows.extend((t, min(t + maxlen, e)) for t in range(int(s), int(e), maxlen))
    return np.asarray(rows, dtype=np.int64).reshape(-1, 2), maxlen


def reshape_and_pad_stacks(x_stack, y_stack, fidx, transform_x=(lambda x: x), transform_y=(lambda y: y), pad_value=0.0,
                           maxlen=None, verbose=False):
    """audio_dataset.py:116-169: (2F, total frames) stacks -> x, y of shape (n_sequences, maxlen, d) filled with
    pad_value beyond each sequence, and mask (n_sequences, maxlen, 1) = 1 on data frames."""
    table, maxlen = sequence_table(fidx, maxlen)
    d = transform_x(x_stack[:, 0:1]).shape[0]
    n = table.shape[0]
    x = np.full((n, maxlen, d), pad_value, dtype=x_stack.dtype)
    y = np.full((n, maxlen, d), pad_value, dtype=y_stack.dtype)
    mask = np.zeros((n, maxlen, 1), dtype=x_stack.dtype)
    for i, (t0, t1) in enumerate(table):
        if verbose:
            print("Sequence %d of %d: t0=%d, t1=%d, duration=%d" % (i + 1, n, t0, t1, t1 - t0))
        x[i, :t1 - t0] = transform_x(x_stack[:, t0:t1]).T
        y[i, :t1 - t0] = transform_y(y_stack[:, t0:t1]).T
        mask[i, :t1 - t0] = 1.0
    return x, y, mask


def clip_x_to_y(x, y, xfidx, yfidx):
    """audio_dataset.py:90-104: keep, per utterance, the first len(y_utt) frames of x; returns a (d, frames_of_y) array."""
    keep = [x[:, xs:xs + (ye - ys)] for (xs, _), (ys, ye) in zip(np.asarray(xfidx), np.asarray(yfidx))]
    return np.concatenate(keep, axis=1)[:, :y.shape[1]]


class AudioDataset:
    """In-memory form: AudioDataset(list of waveforms [, list of target waveforms], params_stft=...).
    Reference form (audio_dataset.py:177-262): AudioDataset(taskfile_input, taskfile_output, datafile=..., params_stft=...,
    downsample=1) with text files listing one wav path per line; the [Re;Im] stacks and the fidx table are cached in
    `datafile` and served from it when it exists (an .npz here; the reference writes HDF5 through h5py)."""

    def __init__(self, x_waveforms, y_waveforms=None, datafile=None, params_stft={"N": 320, "hop": 160, "nch": 1}, downsample=1):
        self.datafile = datafile
        self.x_wavfiles = self.y_wavfiles = None
        if isinstance(x_waveforms, str):          # taskfiles
            from .util import wavread
            if datafile is not None and os.path.isfile(self._cache_name(datafile)):
                self._load_cache(params_stft)
                return
            with open(x_waveforms) as f:
                self.x_wavfiles = [l.strip() for l in f if l.strip()][::downsample]
            with open(y_waveforms) as f:
                self.y_wavfiles = [l.strip() for l in f if l.strip()][::downsample]
            x_waveforms = [wavread(w)[0] for w in self.x_wavfiles]
            y_waveforms = [wavread(w)[0] for w in self.y_wavfiles]
        self._from_waveforms(x_waveforms, y_waveforms, params_stft)
        if datafile is not None:
            self._save_cache()

    @staticmethod
    def _cache_name(datafile):
        return datafile if str(datafile).endswith(".npz") else str(datafile) + ".npz"

    def _save_cache(self):
        y = getattr(self, "y_stack", None)
        np.savez(self._cache_name(self.datafile), x_stack=self.x_stack, y_stack=self.x_stack if y is None else y, fidx=self.fidx,
                 x_wavfiles=np.asarray(self.x_wavfiles or [], dtype=str), y_wavfiles=np.asarray(self.y_wavfiles or [], dtype=str),
                 stft_N=self.params_stft["N"], stft_hop=self.params_stft["hop"])

    def _load_cache(self, params_stft):
        z = np.load(self._cache_name(self.datafile))
        if int(z["stft_N"]) != int(params_stft["N"]) or int(z["stft_hop"]) != int(params_stft["hop"]):
            raise ValueError("datafile '%s' was computed with a different STFT (N=%d, hop=%d)" % (self.datafile, z["stft_N"], z["stft_hop"]))
        self.params_stft = dict(params_stft)
        self.params_stft["window"] = sqrt_hann(self.params_stft["N"])
        dev = torch.device("cuda", torch.cuda.current_device())
        self.x_stack, self.y_stack, self.fidx = z["x_stack"], z["y_stack"], z["fidx"]
        self.x_wavfiles, self.y_wavfiles = list(z["x_wavfiles"]), list(z["y_wavfiles"])
        self.x_stack_dev = torch.as_tensor(self.x_stack, device=dev)
        self.y_stack_dev = torch.as_tensor(self.y_stack, device=dev)
        self.fidx_dev = torch.as_tensor(self.fidx.astype(np.int64), device=dev)

    def _from_waveforms(self, x_waveforms, y_waveforms, params_stft):
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
            self.y_stack_dev, self.y_mag_dev, yfidx_dev = stacks(y_waveforms)
            self.y_stack = self.y_stack_dev.cpu().numpy()
            yfidx = yfidx_dev.cpu().numpy().astype(np.int32)
            if not np.array_equal(yfidx, self.fidx):                              # audio_dataset.py:232-241
                if not np.all(self.fidx[:, 1] - self.fidx[:, 0] >= yfidx[:, 1] - yfidx[:, 0]):
                    raise ValueError("Not all input files have greater than or equal length to all output files!")
                self.x_stack = clip_x_to_y(self.x_stack, self.y_stack, self.fidx, yfidx)
                self.x_stack_dev = torch.as_tensor(self.x_stack, device=dev)
                self.fidx, self.fidx_dev = yfidx, yfidx_dev

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

    def get_data_stacks(self):
        """audio_dataset.py:342-366 (without the HDF5 cache): (x_stack, y_stack, fidx)."""
        return self.x_stack, getattr(self, "y_stack", None), self.fidx

    def get_padded_data_matrix(self, transform_x=(lambda x: x), transform_y=(lambda y: y), pad_value=0.0, maxlen=None):
        """audio_dataset.py:369-383."""
        y_stack = getattr(self, "y_stack", None)
        if y_stack is None:
            y_stack = self.x_stack
        return reshape_and_pad_stacks(self.x_stack, y_stack, self.fidx, transform_x=transform_x, transform_y=transform_y,
                                      pad_value=pad_value, maxlen=maxlen)


def load_data(config, dataset):
    """audio_dataset.py:20-87 for an in-memory AudioDataset: feature maps + pad value from the data config
    ('transform_x', 'transform_y', optional 'maxlen'), returns (x, y, mask) ready for the network."""
    pad = get_mask_value(config)
    return dataset.get_padded_data_matrix(transform_x=data_transform(config.get("transform_x")),
                                          transform_y=data_transform(config.get("transform_y")), pad_value=pad,
                                          maxlen=config.get("maxlen"))
