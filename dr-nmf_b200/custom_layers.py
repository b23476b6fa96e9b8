"""Host-side mirror of the reference's custom_layers.py for the DR-NMF path (same class / function names and
argument meaning), over CUDA tensors and the libdrnmf kernels instead of Keras/Theano.

Reference interfaces mirrored: SimpleDeepRNN (custom_layers.py:104-412), DenseNonNegW (:15-29),
DivideAbyAplusB / divide_A_by_AplusB (:33-56).  What a fixed kernel cannot honour raises instead of silently
computing something else: arbitrary `maps_from_alt` lambdas, activations other than relu, dropout, dense U.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine as _engine


class BuildAltMaps(dict):
    """The `maps_from_alt` object returned by enhance.build_alt: instead of Theano lambdas it records which alt
    parameter feeds which layer (labels_per_k of enhance.py:149-159); the CUDA kernels implement the maps."""

    def __init__(self, labels_per_k, K_layers, output_dim):
        super().__init__(U="build_alt", S="build_alt", W="build_alt", b="build_alt")
        self.labels_per_k, self.K_layers, self.output_dim = labels_per_k, K_layers, output_dim


class SimpleDeepRNN:
    """K_layers-deep network per time step, recurrent over time (custom_layers.py:104).  Only the configuration
    build_unfolded_snmf uses is supported (activation='relu', flag_connect_input_to_layers=True,
    flag_nonnegative=True, maps from build_alt)."""

    def __init__(self, output_dim, init="glorot_uniform", inner_init="orthogonal", activation="tanh",
                 W_regularizer=None, U_regularizer=None, b_regularizer=None, dropout_W=0., dropout_U=0., K_layers=1,
                 alt_params=None, keys_trainable=None, maps_from_alt=None, flag_connect_input_to_layers=False,
                 flag_nonnegative=False, flag_return_all_hidden=False, **kwargs):
        self.units = self.output_dim = int(output_dim)
        self.K_layers = int(K_layers)
        self.alt_params = {} if alt_params is None else alt_params
        self.keys_trainable = list(self.alt_params.keys()) if keys_trainable is None else list(keys_trainable)
        self.maps_from_alt = {} if maps_from_alt is None else maps_from_alt
        self.return_sequences = kwargs.pop("return_sequences", False)
        self.name = kwargs.pop("name", "simple_deep_rnn_1")
        kwargs.pop("input_shape", None)
        if not isinstance(self.maps_from_alt, BuildAltMaps):
            raise NotImplementedError("maps_from_alt must come from drnmf_b200.enhance.build_alt: arbitrary lambdas "
                                      "(custom_layers.py:234-287) cannot be honoured by a fixed CUDA kernel")
        if activation != "relu" or not flag_connect_input_to_layers or not flag_nonnegative:
            raise NotImplementedError("only the DR-NMF configuration is implemented: activation='relu', "
                                      "flag_connect_input_to_layers=True, flag_nonnegative=True (enhance.py:257-266)")
        if dropout_W or dropout_U or W_regularizer or U_regularizer or b_regularizer:
            raise NotImplementedError("dropout / regularizers are off the shipped path")
        self.flag_return_all_hidden = bool(flag_return_all_hidden)
        if not self.return_sequences:
            raise NotImplementedError("return_sequences=False is not used by the DR-NMF path")
        self.built = False
        self._engine = None
        self.log_h0 = None

    def compute_output_shape(self, input_shape):
        units = self.K_layers * self.units if self.flag_return_all_hidden else self.units      # custom_layers.py:176-181
        return (input_shape[0], input_shape[1], units)

    def build(self, input_shape, recon=None):
        """custom_layers.py:187-294.  Like the reference, the numpy arrays in `alt_params` are replaced IN PLACE by
        device-resident variables (CUDA tensors here, Theano shared variables there)."""
        self.input_dim = int(input_shape[2])
        dev = torch.device("cuda", torch.cuda.current_device())
        g = torch.Generator().manual_seed(7654)
        if self.log_h0 is None:   # Keras 'uniform' initializer: U(-0.05, 0.05)
            self.log_h0 = ((torch.rand(self.units, generator=g) - 0.5) * 0.1).to(dev)
        for key in list(self.alt_params):
            v = self.alt_params[key]
            if not torch.is_tensor(v):
                self.alt_params[key] = torch.as_tensor(np.asarray(v, dtype=np.float32)).to(dev)
        self.built = True

    # -- parameters in the engine's layout -----------------------------------------------------------
    def _param_dict(self, k_clean, k_noise):
        lab = self.maps_from_alt.labels_per_k
        K = self.K_layers

        def stack(name):
            labels = lab[name]
            if len(set(labels)) == 1:
                return self.alt_params[labels[0]][None]
            return torch.stack([self.alt_params[l] for l in labels])

        # log_U1 / log_Uk are R x R constants of the a*I + b*11^T form (enhance.py:163-167); verifying that structure
        # means a pass over both matrices on the host, so the (diag, off) pairs are cached until the tensors change
        # (a training step rebuilds the parameter set every iteration)
        key = tuple((id(self.alt_params[n]), getattr(self.alt_params[n], "_version", 0)) for n in ("log_U1", "log_Uk"))
        if getattr(self, "_u_key", None) != key:
            from .engine import structured_u
            self._u_pairs = (structured_u(self.alt_params["log_U1"], "log_U1"), structured_u(self.alt_params["log_Uk"], "log_Uk"))
            self._u_key = key
        return {"log_D": stack("log_D"), "log_alph": stack("log_alph").reshape(len(set(lab["log_alph"])), -1),
                "log_lam1": stack("log_lam1").reshape(-1), "log_U1": self._u_pairs[0], "log_Uk": self._u_pairs[1],
                "log_h0": self.log_h0, "k_clean": k_clean, "k_noise": k_noise}

    @property
    def weights(self):
        """[(name, tensor)] in the reference's naming '{layer}_<key>' (custom_layers.py:205,225); log_h0 first."""
        out = [("%s_log_h0" % self.name, self.log_h0)]
        out += [("%s_%s" % (self.name, k), v) for k, v in self.alt_params.items()]
        return out

    def trainable_keys(self):
        return ["log_h0"] + [k for k in self.alt_params if k in self.keys_trainable]

    def get_config(self):   # custom_layers.py:397-412 (alt_params / maps are not serialisable there either)
        return {"output_dim": self.output_dim, "K_layers": self.K_layers, "activation": "relu",
                "flag_connect_input_to_layers": True}


class DenseNonNegW:
    """Dense layer with kernel exp(kernel) (custom_layers.py:15-29).  It is executed fused with the mask inside the
    model (recon GEMM + DivideAbyAplusB epilogue); kernel layout is Keras' (input_dim, units) = (r, F)."""

    def __init__(self, units, use_bias=False, weights=None, activation=None, name=None, **kwargs):
        if use_bias or activation is not None:
            raise NotImplementedError("DenseNonNegW is used without bias / activation (enhance.py:283,292)")
        self.units, self.name = int(units), name or "dense_non_neg_w"
        self.kernel = None if weights is None else torch.as_tensor(np.asarray(weights[0], dtype=np.float32))

    def set_weights(self, weights):
        self.kernel = torch.as_tensor(np.asarray(weights[0], dtype=np.float32))

    def get_weights(self):
        return [self.kernel.detach().cpu().numpy()]


class DivideAbyAplusB:
    """exp(log(1e-7 + A) - log(1e-7 + A + B)) (custom_layers.py:33-45); fused into the recon GEMM epilogue."""

    def _merge_function(self, inputs):
        raise NotImplementedError("DivideAbyAplusB runs fused inside drnmf_forward (EPI_RECON); build the model with "
                                  "enhance.build_unfolded_snmf")


def divide_A_by_AplusB(inputs, **kwargs):
    """Functional interface (custom_layers.py:48-56): returns the layer object that the model fuses."""
    if len(inputs) != 2:
        raise ValueError("divide_A_by_AplusB takes exactly two inputs")
    return DivideAbyAplusB(**kwargs)


def structured_u_from(alt_params):
    return (_engine.structured_u(alt_params["log_U1"], "log_U1"), _engine.structured_u(alt_params["log_Uk"], "log_Uk"))
