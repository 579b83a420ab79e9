"""Host-side mirror of the reference interface: constructors, state-dict schema, seeded weights, error behaviour.
CPU only: nothing here launches a kernel."""
import json
import os

import pytest
import torch
import torch.nn as nn

from puresound_b200 import recipes, testing
from puresound_b200.nnet.base_nn import SoTaskWrapModule
from puresound_b200.nnet.conv_tasnet import TCN, ConvTasNet
from puresound_b200.nnet.dprnn import DPRNN
from puresound_b200.nnet.lobe.encoder import ConvEncDec, FreeEncDec
from puresound_b200.nnet.lobe.norm import get_norm
from puresound_b200.nnet.lobe.trivial import overlap_geometry

from conftest import GOLDEN
from oracle import separator_ref as R


@pytest.fixture(scope="module")
def pins():
    with open(os.path.join(GOLDEN, "full_size_pins.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("name", ["cfg1", "cfg3", "cfg4", "cfg5_offline", "veve_dprnn_v0_causal"])
def test_seeded_weights_equal_reference(pins, name):
    """torch.manual_seed(0) + perturb(1) reproduces the reference's weights bit for bit (same construction
    order, same keys): parameter count and checksum were recorded from the reference itself."""
    torch.manual_seed(0)
    m = recipes.baseline_config(name).eval()
    testing.perturb_(m, seed=1)
    assert m.overall_parameters == pins[name]["params"]
    assert testing.state_checksum(m.state_dict()) == pytest.approx(pins[name]["state_checksum"], rel=1e-12)


def test_recipe_known_answers():
    """Parameter counts the reference prints (egs/tse/model.py:96-100 minus the 251x192 AAM head; :609-613)."""
    assert recipes.init_model("td_tse_conv_tasnet_v0", verbose=False).overall_parameters == 10108119
    assert recipes.init_model("veve_dprnn_v0_causal", verbose=False).overall_parameters == 723585
    with pytest.raises(NameError):
        recipes.init_model("no_such_model")


def test_state_dict_schema():
    t = TCN(16, 24, 3, 2, emb_dim=4)
    keys = set(t.state_dict())
    for k in ["in_conv.0.weight", "in_conv.1.gamma", "in_conv.1.beta", "in_conv.2.weight", "dconv.0.depthwise.0.weight",
              "dconv.0.depthwise.0.bias", "dconv.0.depthwise.1.weight", "dconv.0.depthwise.2.weight",
              "dconv.0.pointwise.0.weight", "dconv.0.pointwise.1.bias", "dconv.0.pointwise.2.weight", "out_conv.weight", "out_conv.bias"]:
        assert k in keys
    assert t.state_dict()["in_conv.0.weight"].shape == (24, 20, 1)
    e = ConvEncDec(64, "hann", 64, hop_length=16)
    assert set(e.state_dict()) == {"encoder.wsin", "encoder.wcos", "encoder.kernel_sin_inv", "encoder.kernel_cos_inv", "encoder.window_mask"}
    assert isinstance(e.encoder.wsin, nn.Parameter) and e.state_dict()["encoder.kernel_cos_inv"].shape == (64, 1, 64, 1)
    d = DPRNN(16, 12, 16, 2, embed_dim=6, block_with_embed=[0, 1], causal=False)
    keys = set(d.state_dict())
    assert "intra_rnn.0.weight_ih_l0_reverse" in keys and "input_film.1.cond_scale.weight" in keys and "output_fc.1.bias" in keys
    assert d.input_film[0] is None


def test_error_conventions():
    with pytest.raises(NameError):
        get_norm("xLN")
    with pytest.raises(AssertionError):  # lobe/cnn.py:40-44
        TCN(16, 24, 3, 1, causal=True, tcn_norm="cLN", dconv_norm="gGN")
    with pytest.raises(AssertionError):  # conv_tasnet.py:277
        ConvTasNet(16, 0, per_tcn_stack=3, tcn_with_embed=[0, 0])
    with pytest.raises(NameError):  # conv_tasnet.py:275
        ConvTasNet(16, 0, tcn_layer="weird")
    with pytest.raises(TypeError):  # encoder.py:338-339
        ConvEncDec(64, "hann", 32)
    m = SoTaskWrapModule(FreeEncDec(32, 16, 16), ConvTasNet(16, 0, tcn_dim=8, per_tcn_stack=1, repeat_tcn=1, tcn_with_embed=[0]), verbose=False)
    assert m.task == 0
    with pytest.raises(Exception):  # no CPU fallback: must fail loudly, never compute on the host
        m.inference(torch.zeros(1, 400))


@pytest.mark.parametrize("T,K", [(103, 10), (9999, 100), (100, 20), (57, 8), (1, 4)])
def test_overlap_geometry_matches_oracle(T, K):
    seg, rest = R.split_overlap(torch.zeros(1, 2, T), K)
    r, S = overlap_geometry(T, K)
    assert (r, S) == (rest, seg.shape[1])


def test_get_args_roundtrip():
    a = ConvTasNet(32, 8, True, tcn_dim=16, per_tcn_stack=2, repeat_tcn=1, tcn_with_embed=[1, 0]).get_args
    assert ConvTasNet(**a).get_args == a


def test_mel_front_end_schema_and_filters():
    """FbankEnc / ConvMelSpectrogram (lobe/encoder.py:186-272,459-507): state-dict keys, buffers vs parameters, the Slaney mel
    filters' defining properties; against the live reference when it is present (authoring container)."""
    import sys

    from puresound_b200.nnet.lobe.encoder import FbankEnc, mel_filterbank

    fixed, train = FbankEnc(trainable=False, n_banks=80), FbankEnc(trainable=True, n_banks=40)
    assert list(fixed.state_dict()) == ["encoder.wsin", "encoder.wcos", "encoder.window_mask", "encoder.filterbank", "encoder.inv_filterbank"]
    assert len(list(fixed.parameters())) == 0 and len(list(train.parameters())) == 4
    assert fixed.state_dict()["encoder.filterbank"].shape == (257, 80) and fixed.state_dict()["encoder.inv_filterbank"].shape == (80, 257)
    fb = mel_filterbank(16000, 512, 80)
    assert fb.shape == (80, 257) and fb.dtype == torch.float32 and (fb >= 0).all() and (fb.max(1).values > 0).all()
    peaks = fb.argmax(1)
    assert (peaks[1:] >= peaks[:-1]).all() and fb[:, 0].abs().max() == 0  # centres rise with the band; DC belongs to no band
    with pytest.raises(NotImplementedError):
        FbankEnc(output_format="MagPhase")
    if os.path.isdir("/root/reference/puresound"):
        sys.path.insert(0, "/root/reference")
        try:
            from puresound.nnet.lobe.encoder import FbankEnc as RefFbank
        finally:
            sys.path.remove("/root/reference")
        for ours, kw in ((fixed, dict(trainable=False, n_banks=80)), (train, dict(trainable=True, n_banks=40))):
            ref = RefFbank(**kw).state_dict()
            assert list(ref) == list(ours.state_dict()) and all(torch.equal(ref[k], ours.state_dict()[k]) for k in ref)


def test_spec_augment_draws_like_torchaudio():
    """SpecAugment's host-side band draw consumes the global generator exactly like torchaudio.functional.mask_along_axis
    (two torch.rand(1) per masked axis), so the same seed gives the reference's band."""
    from puresound_b200.nnet.lobe.trivial import SpecAugment

    torch.manual_seed(11)
    v = torch.rand(1) * 10
    v0 = torch.rand(1) * (80 - v)
    nxt = torch.rand(1)
    torch.manual_seed(11)
    s, e = SpecAugment._draw(10, 80)
    assert (s, e) == (int(v0.long()), int(v0.long()) + int(v.long())) and torch.equal(torch.rand(1), nxt)
    torch.manual_seed(11)
    assert SpecAugment._draw(0, 80) == (0, 0) and torch.equal(torch.rand(1), v / 10)  # no mask, no draw


def test_bench_clock_sampler_window_and_roofline_bytes():
    """bench.py host logic that needs no GPU: only the clock samples received inside the timed window are summarised (the
    sampler runs from before the warm-up), and a GEMM group's algorithmic bytes include the residual / mask operand reads of
    its launches."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    clk = bench.ClockSampler(0)
    row = lambda mhz, cap: ["0", str(mhz), "1965", "990.0", "Not Active", "Not Active", "Not Active", "Active" if cap else "Not Active"]
    clk.rows = [(1.0, row(1965, False)), (2.0, row(1550, True)), (2.5, row(1560, True)), (3.5, row(1965, False))]
    clk.t_open, clk.t_close = 1.5, 3.0
    s = clk.summary()
    assert s["samples"] == 2 and s["sm_mhz"] == 1555.0 and s["sm_max_mhz"] == 1965.0 and s["reasons"] == ["sw_power_cap"]
    assert "extended" not in s
    clk.extended = True
    assert "extended" in clk.summary()

    class Ev:  # stands in for a CUDA event pair: elapsed_time in ms
        def __init__(self, t):
            self.t = t

        def elapsed_time(self, other):
            return other.t - self.t

    rows, M, K = 1000, 128, 256
    events = [("gemm", Ev(0.0), Ev(0.5), (rows, M, K), 4 * rows * M), ("gemm", Ev(1.0), Ev(1.5), (rows, M, K), 0)]
    (g,) = bench.kernel_rooflines(events, 2.0, 6549.8, 1392.7, "measured")
    assert g["launches_timed"] == 2 and abs(g["avg_launch_ms"] - 0.5) < 1e-9
    assert g["algorithmic_bytes"] == 4.0 * (rows * K + rows * M + M * K) + 2 * rows * M  # mean residual bytes of the two launches
