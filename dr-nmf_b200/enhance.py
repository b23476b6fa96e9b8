"""Host-side mirror of the reference's enhance.py for the DR-NMF path: build_alt, build_unfolded_snmf, the predict +
reconstruct loop and the `-c <model yaml> -d <data yaml>` entry point, over libdrnmf instead of Keras/Theano.

Reference: enhance.py:139-206 (build_alt), :209-317 (build_unfolded_snmf), :1183-1203 (predict/reconstruct),
:459-538 (CLI; the model kind is inferred from the config FILE NAME).  CHiME2 audio is not available offline, so
`main` runs the same pipeline on the synthetic workload of drnmf_b200.synth (--synthetic, default).
"""
from __future__ import annotations

import getopt
import os
import sys

import numpy as np
import torch

from . import engine as _engine
from . import synth as _synth
from .custom_layers import BuildAltMaps, DenseNonNegW, SimpleDeepRNN, divide_A_by_AplusB


def build_alt(output_dim, K_layers, params, params_untied=[]):
    """enhance.py:139-206.  Returns (alt_params, maps_from_alt): alt_params are the same float32 numpy arrays the
    reference creates (log(1e-7 + .) parameterisation, per-layer copies for untied names); maps_from_alt describes
    which parameter feeds which layer (the maps themselves are the CUDA kernels of drnmf_set_params)."""
    e32 = np.float32(1e-7)
    alt_params = {"log_D": np.log(1e-7 + params["W"]).astype(np.float32),
                  "log_U1": np.log(e32 + params["U1"]), "log_Uk": np.log(e32 + params["Uk"]),
                  "log_alph": np.log(e32 + params["alph"]), "log_lam1": np.log(e32 + params["lam1"])}
    labels_per_k = {}
    for name in ["log_D", "log_alph", "log_lam1"]:
        if name in params_untied:
            labels_per_k[name] = [name + ("_%d" % k) for k in range(K_layers)]
            p = alt_params.pop(name)
            for k in range(K_layers):
                alt_params[name + ("_%d" % k)] = p
        else:
            labels_per_k[name] = [name] * K_layers
    return alt_params, BuildAltMaps(labels_per_k, K_layers, output_dim)


class UnfoldedSNMFModel:
    """What build_unfolded_snmf returns: Masking -> SimpleDeepRNN -> H_clean/H_noise -> DenseNonNegW x2 ->
    divide_A_by_AplusB (enhance.py:251-305), with the Keras Model surface the reference uses."""

    def __init__(self, rnn, clean_est, noise_est, merge, mask_value, maxseq, input_dim, square):
        self.rnn, self.clean_est, self.noise_est, self.merge = rnn, clean_est, noise_est, merge
        self.mask_value, self.maxseq, self.input_dim, self.square = mask_value, maxseq, input_dim, square
        self.layers = ["input", "masking", rnn, "H_clean", clean_est, "H_noise", noise_est, merge]
        self._eng = None
        self._dirty = True

    # -- weights ---------------------------------------------------------------------------------------
    def get_weights(self):
        w = [t.detach().cpu().numpy() for _, t in self.rnn.weights]
        return w + self.clean_est.get_weights() + self.noise_est.get_weights()

    def weight_names(self):
        return [n for n, _ in self.rnn.weights] + ["clean_est/kernel", "noise_est/kernel"]

    def set_weights(self, weights):
        names = self.weight_names()
        if len(weights) != len(names):
            raise ValueError("expected %d weight arrays, got %d" % (len(names), len(weights)))
        dev = torch.device("cuda", torch.cuda.current_device())
        it = iter(weights)
        self.rnn.log_h0 = torch.as_tensor(np.asarray(next(it), np.float32)).to(dev)
        for k in list(self.rnn.alt_params):
            self.rnn.alt_params[k] = torch.as_tensor(np.asarray(next(it), np.float32)).to(dev)
        self.clean_est.set_weights([next(it)])
        self.noise_est.set_weights([next(it)])
        self._dirty = True

    def save_weights(self, path):
        np.savez(path, **{n.replace("/", "__"): w for n, w in zip(self.weight_names(), self.get_weights())})

    def load_weights(self, path):
        z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
        self.set_weights([z[n.replace("/", "__")] for n in self.weight_names()])

    # -- inference -------------------------------------------------------------------------------------
    def _engine_ready(self):
        if self._eng is None:
            self._eng = _engine.DrnmfEngine(self.input_dim, self.rnn.units, self.rnn.K_layers, square_irm=self.square)
        if self._dirty:
            self._eng.set_params(self.rnn._param_dict(self.clean_est.kernel, self.noise_est.kernel))
            self._dirty = False
        return self._eng

    def predict_on_batch(self, x, return_hidden=False):
        """enhance.py:1191-1193: x (B, T, F) numpy (or CUDA tensor) padded with mask_value -> irm (B, T, F)."""
        eng = self._engine_ready()
        xt = x if torch.is_tensor(x) else torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32))
        xt = xt.to(eng.device, torch.float32)
        H, irm = eng.forward(xt, self.mask_value, want_H=return_hidden)
        if return_hidden and self.rnn.flag_return_all_hidden:      # the layer's output is the concatenation of all layers
            H = eng.forward_all_hidden(xt, self.mask_value)
        if torch.is_tensor(x):
            return (irm, H) if return_hidden else irm
        irm = irm.cpu().numpy()
        return (irm, H.cpu().numpy()) if return_hidden else irm

    def fit_pretrain(self, x, lam1, batch_size=32, epochs=1, validation_data=None, learning_rate=1e-3, clipnorm=0.0,
                     decay=0.0, patience=50, savefile=None, verbose=0):
        """Optional pretraining with the SNMF cost (enhance.py:1024-1036, 1088-1120): model_pretrain.fit(x, [x, x], ...)
        with losses ['mse' of x_recon = S^ + N^, l1 of the hidden output] and weights [0.5, lam1 * 2r / F]."""
        from .training import Trainer
        tr = Trainer(self, learning_rate=learning_rate, clipnorm=clipnorm, decay=decay, loss="snmf_cost", lam1=lam1)
        vd = None if validation_data is None else (validation_data, validation_data)
        try:
            return tr.fit(x, x, batch_size=batch_size, epochs=epochs, validation_data=vd, patience=patience, savefile=savefile,
                          verbose=verbose)
        finally:
            self._engine_ready().set_training_loss("mse_of_masked")

    def fit(self, x, y, sample_weight=None, batch_size=32, epochs=1, validation_data=None, learning_rate=1e-3,
            clipnorm=0.0, decay=0.0, patience=50, savefile=None, verbose=0):
        """enhance.py:1152-1157 (loss 'mse' on x*irm with temporal sample weights, Adam): see training.Trainer.
        sample_weight is accepted for signature compatibility; the frame mask is derived from the -1 padding."""
        from .training import Trainer
        if getattr(self, "_trainer", None) is None:
            self._trainer = Trainer(self, learning_rate=learning_rate, clipnorm=clipnorm, decay=decay)
        vd = None if validation_data is None else (validation_data[0], validation_data[1])
        return self._trainer.fit(x, y, batch_size=batch_size, epochs=epochs, validation_data=vd, patience=patience,
                                 savefile=savefile, verbose=verbose)


def build_unfolded_snmf(params_unfolded_snmf):
    """enhance.py:209-317 with the same params dict keys: input_dim, hidden_dim, output_dim, mask_value, maxseq,
    K_layers, W, alph, lam1 [, untie_alph, params_untied, params_trainable, transform_before_irm]."""
    p = params_unfolded_snmf
    input_dim, hidden_dim, output_dim = int(p["input_dim"]), int(p["hidden_dim"]), int(p["output_dim"])
    K_layers, W_noisy = int(p["K_layers"]), np.float32(p["W"])
    params_const = {"W": W_noisy, "U1": np.eye(hidden_dim).astype(np.float32),
                    "Uk": np.zeros((hidden_dim, hidden_dim)).astype(np.float32), "alph": np.float32(p["alph"]),
                    "lam1": np.float32(p["lam1"])}
    if p.get("untie_alph"):
        params_const["alph"] = params_const["alph"] * np.ones((hidden_dim,), dtype=np.float32)
    params_untied = p.get("params_untied", [])
    alt_params, maps_from_alt = build_alt(hidden_dim, K_layers, params_const, params_untied=params_untied)
    if "params_trainable" not in p:
        # the reference only assigns keys_trainable inside this branch and then uses it unconditionally
        # (enhance.py:239-263): a missing key is a NameError there; here it is an explicit error.
        raise KeyError("params_unfolded_snmf['params_trainable'] is required (enhance.py:239-248)")
    keys_trainable = []
    for name in p["params_trainable"]:
        keys_trainable += [name + ("_%d" % k) for k in range(K_layers)] if name in params_untied else [name]
    rnn = SimpleDeepRNN(hidden_dim, input_shape=(p["maxseq"], input_dim), return_sequences=True, activation="relu",
                        K_layers=K_layers, alt_params=alt_params, keys_trainable=keys_trainable,
                        maps_from_alt=maps_from_alt, flag_connect_input_to_layers=True, flag_nonnegative=True,
                        flag_return_all_hidden=bool(p.get("flag_return_all_hidden", False)))
    rnn.build((None, p["maxseq"], input_dim))
    r = hidden_dim // 2
    log_W_clean = np.log(1e-7 + W_noisy[:, :r])
    log_W_noise = np.log(1e-7 + W_noisy[:, r:])
    clean_est = DenseNonNegW(output_dim, use_bias=False, weights=[log_W_clean.T], name="clean_est")
    noise_est = DenseNonNegW(output_dim, use_bias=False, weights=[log_W_noise.T], name="noise_est")
    square = False
    if "transform_before_irm" in p:
        if p["transform_before_irm"] == "square":
            square = True
        else:   # the reference constructs this error and drops it (enhance.py:302); here it is raised
            raise ValueError("Unknown 'transform_before_irm' of '%s'" % (p["transform_before_irm"]))
    merge = divide_A_by_AplusB([clean_est, noise_est])
    return UnfoldedSNMFModel(rnn, clean_est, noise_est, merge, float(p["mask_value"]), p["maxseq"], input_dim, square)


# ------------------------------------------------------------------------------------------------------
def enhance_batch(model, noisy_list, N, hop, batch_size=250):
    """enhance.py:1186-1203: STFT -> padded magnitudes -> predict in slabs of `batch_size` -> mask -> iSTFT.
    noisy_list: list of 1-D float32 waveforms.  Returns the list of enhanced waveforms (numpy)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    lens = [len(x) for x in noisy_list]
    offs = np.concatenate([[0], np.cumsum(lens)])[:-1]
    audio = torch.as_tensor(np.concatenate(noisy_list).astype(np.float32), device=dev)
    stack, mag, fidx = _engine.stft_mag(audio, list(offs), lens, N, hop)
    F = N // 2 + 1
    fi = fidx.cpu().numpy()
    counts = fi[:, 1] - fi[:, 0]
    maxseq = int(counts.max())
    n = len(noisy_list)
    x = torch.full((n, maxseq, F), model.mask_value, dtype=torch.float32, device=dev)
    for j in range(n):
        x[j, :counts[j]] = mag[fi[j, 0]:fi[j, 1]]
    irm_frames = torch.empty((int(fi[-1, 1]), F), dtype=torch.float32, device=dev)
    start = 0
    while start < n:                                   # slabs of 250 utterances (enhance.py:1189-1193)
        irm = model.predict_on_batch(x[start:start + batch_size])
        for j in range(start, min(n, start + batch_size)):
            irm_frames[fi[j, 0]:fi[j, 1]] = irm[j - start, :counts[j]]
        start += batch_size
    ys = _engine.mask_istft(stack, irm_frames, fidx, N, hop)
    return [y.cpu().numpy()[:L] for y, L in zip(ys, lens)]


def main(argv):
    """`enhance.py -c <model yaml> -d <data yaml>` (enhance.py:459-488); model kind from the config file name
    (:530-538).  Extra flags: --synthetic N (number of synthetic utterances, default 8), --seconds S."""
    import yaml
    configfile = datafile = ""
    n_synth, seconds = 8, 3.0
    opts, _ = getopt.getopt(argv, "hc:d:", ["cfile=", "dfile=", "synthetic=", "seconds="])
    for opt, arg in opts:
        if opt == "-h":
            print("enhance.py -c <model yaml> -d <data yaml> [--synthetic N] [--seconds S]")
            return 0
        elif opt in ("-c", "--cfile"):
            configfile = arg
        elif opt in ("-d", "--dfile"):
            datafile = arg
        elif opt == "--synthetic":
            n_synth = int(arg)
        elif opt == "--seconds":
            seconds = float(arg)
    params = yaml.safe_load(open(configfile)) if configfile else {}
    data = yaml.safe_load(open(datafile)) if datafile else {"params_stft": {"N": 512, "hop": 128, "nch": 1}}
    if configfile and "unfolded_snmf" not in os.path.basename(configfile):
        raise NotImplementedError("only 'unfolded_snmf' configs belong to the hot path (the 'snmf' and 'lstm' model "
                                  "kinds of enhance.py:530-538 are out of scope, SURVEY section 2)")
    N, hop = int(data["params_stft"]["N"]), int(data["params_stft"]["hop"])
    F = N // 2 + 1
    r = int(params.get("r", 100))
    K_layers = int(params.get("K_layers", 2))
    alph, lam1 = float(params.get("alph", 50.0)), float(params.get("lam1", 1.0))
    # weight_initialization: 'snmf' needs CHiME2 training frames; offline we use the synthetic template dictionary
    W = _synth.dictionary(F, 2 * r)
    build = {"input_dim": F, "hidden_dim": 2 * r, "output_dim": F, "mask_value": -1.0, "maxseq": data.get("maxlen", 500),
             "K_layers": K_layers, "W": W, "alph": alph, "lam1": lam1,
             "params_untied": params.get("params_untied", ["log_D", "log_alph"]),
             "params_trainable": params.get("params_trainable", ["log_D", "log_alph"])}
    for k in ("untie_alph", "transform_before_irm"):
        if k in params:
            build[k] = params[k]
    model = build_unfolded_snmf(build)
    pairs = [_synth.utterance(i, seconds=seconds) for i in range(n_synth)]
    enhanced = enhance_batch(model, [n for n, _ in pairs], N, hop)
    from . import scoring
    sdr_in = np.mean([scoring.sdr_db(n, c) for n, c in pairs])
    sdr_out = np.mean([scoring.sdr_db(e, c) for e, (_, c) in zip(enhanced, pairs)])
    print("DR-NMF (K=%d, r=%d, F=%d): %d synthetic utterances, mean SDR in %.2f dB -> out %.2f dB (untrained dictionary)"
          % (K_layers, r, F, n_synth, sdr_in, sdr_out))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
