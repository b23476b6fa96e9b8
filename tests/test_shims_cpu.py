"""CPU tests of the host-side mirrors of the reference interfaces (no GPU work)."""
import numpy as np
import pytest

import oracle as O
from drnmf_b200 import custom_layers, enhance, synth


def _const(W, alph, lam1, R, untie=False):
    p = {"W": np.float32(W), "U1": np.eye(R).astype(np.float32), "Uk": np.zeros((R, R)).astype(np.float32),
         "alph": np.float32(alph), "lam1": np.float32(lam1)}
    if untie:
        p["alph"] = p["alph"] * np.ones((R,), dtype=np.float32)
    return p


def test_build_alt_matches_reference_parameterisation():
    """oracle.alt_params_init is pinned bit-for-bit against the reference's build_alt (oracle/pin_reference.py);
    the shim must produce the same arrays under the reference's key names."""
    F, R, K = 12, 6, 3
    W = synth.dictionary(F, R)
    alt, maps = enhance.build_alt(R, K, _const(W, 50., 1., R), params_untied=["log_D", "log_alph"])
    ref = O.alt_params_init(W, 50., 1., K)
    assert sorted(alt) == sorted(["log_D_%d" % k for k in range(K)] + ["log_alph_%d" % k for k in range(K)]
                                 + ["log_lam1", "log_U1", "log_Uk"])
    for k in range(K):
        np.testing.assert_array_equal(alt["log_D_%d" % k], ref["log_D"][k])
        np.testing.assert_array_equal(alt["log_alph_%d" % k], ref["log_alph"][k])
    np.testing.assert_array_equal(alt["log_lam1"], ref["log_lam1"][0])
    np.testing.assert_array_equal(alt["log_U1"], ref["log_U1"])
    np.testing.assert_array_equal(alt["log_Uk"], ref["log_Uk"])
    assert maps.labels_per_k["log_lam1"] == ["log_lam1"] * K and isinstance(maps, custom_layers.BuildAltMaps)
    # tied: one shared array
    alt_t, maps_t = enhance.build_alt(R, K, _const(W, 50., 1., R), params_untied=[])
    assert sorted(alt_t) == ["log_D", "log_U1", "log_Uk", "log_alph", "log_lam1"]
    assert maps_t.labels_per_k["log_D"] == ["log_D"] * K


def test_trainable_weight_count_identity():
    """notebook :121-126: K*F*2r + K + 2r counts log_D_k, log_alph_k and log_h0."""
    F, r, K = 257, 100, 2
    alt, _ = enhance.build_alt(2 * r, K, _const(np.ones((F, 2 * r)), 50., 1., 2 * r), ["log_D", "log_alph"])
    n = sum(int(np.size(alt["log_D_%d" % k])) + int(np.size(alt["log_alph_%d" % k])) for k in range(K)) + 2 * r
    assert n == 103002 == O.param_count_notebook(F, r, K)


def test_layer_rejects_what_the_kernel_cannot_do():
    with pytest.raises(NotImplementedError):
        custom_layers.SimpleDeepRNN(8, activation="relu", K_layers=2, alt_params={}, maps_from_alt={"U": [lambda a: a]},
                                    flag_connect_input_to_layers=True, flag_nonnegative=True, return_sequences=True)
    _, maps = enhance.build_alt(4, 2, _const(np.ones((3, 4)), 5., 1., 4), [])
    with pytest.raises(NotImplementedError):
        custom_layers.SimpleDeepRNN(4, activation="tanh", K_layers=2, alt_params={}, maps_from_alt=maps,
                                    flag_connect_input_to_layers=True, flag_nonnegative=True, return_sequences=True)
    with pytest.raises(NotImplementedError):
        custom_layers.DenseNonNegW(5, use_bias=True)
    with pytest.raises(ValueError):
        custom_layers.divide_A_by_AplusB([1, 2, 3])


# ---- data formats either side of the path: pinned against the reference's own code -----------------------
@pytest.mark.parametrize("tag,maxlen", [("full", None), ("m5", 5), ("m4", 4), ("big", 100)])
def test_reshape_and_pad_stacks_golden(golden_dir, tag, maxlen):
    """tests/golden/dataset.npz holds the outputs of the reference's reshape_and_pad_stacks (audio_dataset.py:116-169);
    the oracle restatement and the product mirror must both reproduce them exactly, chunking included."""
    import os
    from drnmf_b200 import audio_dataset as ad
    g = np.load(os.path.join(golden_dir, "dataset.npz"))
    for impl, mag in ((O.reshape_and_pad_stacks, O.data_transform("mag")), (ad.reshape_and_pad_stacks, ad.data_transform("mag"))):
        x, y, m = impl(g["x_stack"].copy(), g["y_stack"].copy(), g["fidx"], transform_x=mag, transform_y=mag, pad_value=-1.0,
                       maxlen=maxlen)
        np.testing.assert_array_equal(x, g[tag + "_x"])
        np.testing.assert_array_equal(y, g[tag + "_y"])
        np.testing.assert_array_equal(m, g[tag + "_mask"])
    # the pad value is what Masking(-1) keys on, the mask marks exactly the data frames
    assert np.all(g[tag + "_x"][g[tag + "_mask"][..., 0] == 0] == -1.0)
    assert int(g[tag + "_mask"].sum()) == int(g["fidx"][-1, 1])


def test_clip_mask_value_and_savefile_golden(golden_dir):
    import os
    from drnmf_b200 import audio_dataset as ad, snmf
    g = np.load(os.path.join(golden_dir, "dataset.npz"))
    for impl in (O.clip_x_to_y, ad.clip_x_to_y):
        np.testing.assert_array_equal(impl(g["clip_x"].copy(), g["y_stack"], g["clip_xfidx"], g["fidx"]), g["clip_out"])
    for cfg, want in (({"transform_x": "mag", "transform_y": "mag"}, -1.0), ({"transform_x": "x", "transform_y": "logmag"}, -1.0),
                      ({"transform_x": "x", "transform_y": "y"}, 0.0)):
        assert O.get_mask_value(cfg) == want and ad.get_mask_value(cfg) == want
    prm = {"cf": "ed", "sparsity": np.float32(5.0), "max_iter": 200., "conv_eps": 1e-4, "display": 0., "random_seed": 2016.,
           "r": np.int64(100)}
    ref_name = str(g["savefile_name"])                        # produced by the reference's get_snmf_savefile
    assert O.snmf_savefile_stem(prm, "dicts/") + ".hkl" == ref_name
    assert snmf.get_snmf_savefile(prm, "dicts/") == ref_name[:-4] + ".npz"


def test_scoring_snr_and_table():
    from drnmf_b200 import scoring
    rng = np.random.default_rng(3)
    ref = rng.standard_normal(4000)
    est = ref + 0.1 * rng.standard_normal(4000)
    assert abs(scoring.snr_db(est, ref) - O.snr_db(est, ref)) < 1e-12
    assert 19.0 < scoring.snr_db(est, ref) < 21.0
    row, labels = scoring.compute_scores(est[:3900], ref)
    assert labels[:2] == ["SDR", "SNR"] and len(row) == len(labels) == 6
    assert abs(row[0] - O.sdr_db(est[:3900], ref)) < 1e-9 and np.isfinite(row[2:4]).all() and np.isnan(row[4:]).all()


def test_wav_io_int16_convention(tmp_path):
    """util.py:29-45: float32 -> int16 (x * 32767, peak-normalised only when clipping) -> float32 / 32768."""
    from drnmf_b200 import util, scoring
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((1, 800)) * 0.2).astype(np.float32)
    f = str(tmp_path / "a.wav")
    util.wavwrite(f, 16000, x)
    y = util.wavread(f)
    assert y.shape == x.shape and y.dtype == np.float32
    np.testing.assert_allclose(y, np.int16(x * 32767.0).astype(np.float32) / 32768.0)
    np.testing.assert_allclose(y[0], scoring.wav_quantize(x[0]))
    big = (x * 10).astype(np.float32)                       # clips: normalised by its peak first
    util.wavwrite(f, 16000, big)
    assert abs(np.abs(util.wavread(f)).max() - 32767.0 / 32768.0) < 1e-4
    assert util.wavread([f]).shape == (1, 800)


def test_segsnr_and_per_snr_aggregation():
    """score_audio.m:209-212 / print_scores.py:84-114."""
    from drnmf_b200 import scoring
    rng = np.random.default_rng(1)
    fs = 16000
    ref = rng.standard_normal(fs)
    est = ref + 0.1 * rng.standard_normal(fs)
    loc, glo = scoring.snrseg(est, ref, fs)
    assert abs(glo - scoring.snr_db(est, ref)) < 1e-9            # whole frames here: global segSNR = raw SNR
    assert abs(loc - 20.0) < 0.5 and abs(glo - 20.0) < 0.3
    loc2, _ = scoring.snrseg(ref, ref, fs)                       # perfect estimate: clipped at 35 dB per frame
    assert loc2 == 35.0
    loc3, _ = scoring.snrseg(-5 * ref, ref, fs)                  # terrible estimate: clipped at -10 dB
    assert loc3 == -10.0
    row, labels = scoring.compute_scores(est, ref, fs)
    assert labels[2] == "SegSNR local" and abs(row[2] - loc) < 1e-12 and np.isnan(row[4]) and np.isnan(row[5])
    per = {"m6dB": np.array([[1.0, 2.0, 0, 0, 0, 0], [3.0, 4.0, 0, 0, 0, 0]]), "9dB": np.array([[10.0, 20.0, 0, 0, 0, 0]])}
    agg, latex = scoring.aggregate_scores(per, scores_to_print=("SDR", "SNR"))
    assert agg["SDR"]["per_snr"] == {"m6dB": 2.0, "9dB": 10.0} and abs(agg["SDR"]["all"] - 14.0 / 3) < 1e-12
    assert latex == "2.00 & 10.00 & 4.67 & 3.00 & 20.00 & 8.67 \\\\"
