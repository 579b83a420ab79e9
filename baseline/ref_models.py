"""Builders for the reference arm of bench.py (`--impl reference`) and its `cpu_baseline` leg.

Imports the UNMODIFIED reference (mcw519/PureSound) from ``baseline/_ref`` - the git-ignored directory
``__graft_entry__.build()`` fills with ``pip install --no-index --no-deps --target baseline/_ref`` of the reference
(plus its two recipe registries ``egs/{tse,ns}/model.py``, which the reference's setup.py does not package) - and
constructs the benchmark workloads from the reference's OWN classes, exactly as ``tests/golden/make_golden.py`` does
for the golden vectors: ``torch.manual_seed(0)`` init + ``testing.perturb_(seed=1)``, i.e. the same weights as the
B200 arm (our modules reproduce the reference's seeded init bit for bit, tests/test_host_logic.py).

Nothing of ours sits on that path: the timed call is the reference's ``SoTaskWrapModule.inference``
(puresound/nnet/base_nn.py:690-722) on CPU tensors.
"""
from __future__ import annotations

import importlib.util
import io
import os
import sys
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "puresound", "nnet"))


def _quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _recipes(kind: str):
    path = os.path.join(REF, "egs", kind, "model.py")
    spec = importlib.util.spec_from_file_location(f"ref_{kind}_model", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build(workload: str):
    """The reference model of a bench workload (reference classes only), seeded like the B200 arm; eval mode."""
    if not available():
        raise FileNotFoundError("baseline/_ref is missing: run __graft_entry__.build() where /root/reference exists")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import torch
    import torch.nn as nn
    from puresound.nnet.base_nn import SoTaskWrapModule
    from puresound.nnet.conv_tasnet import TCN, ConvTasNet, GatedTCN
    from puresound.nnet.dprnn import DPRNN
    from puresound.nnet.lobe.encoder import ConvEncDec, FreeEncDec
    from puresound.nnet.lobe.pooling import AttentiveStatisticsPooling
    from puresound.nnet.lobe.trivial import Magnitude

    sys.path.insert(0, os.path.dirname(HERE))
    from puresound_b200 import testing

    def tasnet(width):
        return ConvTasNet(width, 0, False, tcn_kernel=3, tcn_dim=512, repeat_tcn=3, tcn_dilated_basic=2, per_tcn_stack=8,
                          tcn_with_embed=[0] * 8, tcn_norm="gLN", dconv_norm="gGN", causal=False, tcn_layer="normal")

    torch.manual_seed(0)
    if workload in ("cfg1", "cfg2"):
        m = _quiet(SoTaskWrapModule, encoder=FreeEncDec(32, 512, 16), masker=tasnet(512), mask_constraint="ReLU", verbose=False)
    elif workload == "cfg1b":
        m = _quiet(SoTaskWrapModule, encoder=FreeEncDec(32, 128, 16), masker=tasnet(128), mask_constraint="ReLU", verbose=False)
    elif workload == "cfg3":
        m = _quiet(SoTaskWrapModule, encoder=FreeEncDec(32, 128, 16, output_active=True),
                   masker=DPRNN(128, 128, 128, n_blocks=6, seg_size=100, seg_overlap=True, causal=False), mask_constraint="ReLU", verbose=False)
    elif workload in ("cfg4", "cfg4_gated"):
        if workload == "cfg4":
            masker = ConvTasNet(512, 192, True, tcn_dim=256, repeat_tcn=3, per_tcn_stack=8, tcn_with_embed=[1, 0, 0, 0, 0, 0, 0, 0])
            spk = [TCN(256, 256, 3, dilation=2 ** i) for i in range(5)]
        else:
            masker = ConvTasNet(512, 192, True, tcn_layer="gated", tcn_dim=256, repeat_tcn=3, per_tcn_stack=5, tcn_with_embed=[1, 0, 0, 0, 0])
            spk = [GatedTCN(256, 128, 3, dilation=2 ** i, causal=False, tcn_norm="gLN") for i in range(5)]
        m = _quiet(SoTaskWrapModule, encoder=ConvEncDec(512, "hann", 512, hop_length=128, trainable=True, output_format="Complex"),
                   masker=masker, speaker_net=nn.ModuleList([Magnitude(drop_first=False)] + spk + [AttentiveStatisticsPooling(256, 128), nn.Conv1d(512, 192, 1, bias=False)]),
                   mask_constraint="linear", drop_first_bin=True, verbose=False)
    elif workload == "cfg5":
        # the reference has no streaming Conv-TasNet: its CPU arithmetic for this model is the OFFLINE causal forward
        m = _quiet(SoTaskWrapModule, encoder=FreeEncDec(320, 512, 160),
                   masker=ConvTasNet(512, 0, False, tcn_dim=512, per_tcn_stack=8, repeat_tcn=3, tcn_with_embed=[0] * 8, tcn_norm="cLN",
                                     dconv_norm="cLN", causal=True), mask_constraint="ReLU", verbose=False)
    elif workload.startswith("ns_"):
        m = _quiet(_recipes("ns").init_model, workload, None, verbose=False)
    else:
        m = _quiet(_recipes("tse").init_model, workload, None, None, verbose=False)
    m = m.eval()
    testing.perturb_(m, seed=1)
    return m
