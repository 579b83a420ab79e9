"""Generate the golden vectors under tests/golden/ from the REFERENCE itself.

Run in the authoring container only (it imports /root/reference, which does not
exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Small cases (``small_*.pt``): reference modules with tiny dimensions, seeded
'perturbed' weights (puresound_b200.testing.perturb_), seeded inputs; each file
stores cfg + state_dict + inputs + the reference's outputs.

Full-size pins (``full_size_pins.json``): the BASELINE.json configurations
built under torch.manual_seed(0) (+ perturb seed 1), run through the reference's
``SoTaskWrapModule.inference`` on the seeded synthetic inputs; stores a weight
checksum and a strided sub-sample of the output so that a GPU-box test can
rebuild identical weights from the seed and compare against the reference's own
numbers without the reference being present.
"""
from __future__ import annotations

import importlib.util
import io
import json
import os
import sys
from contextlib import redirect_stdout

REF = os.environ.get("PURESOUND_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

from puresound.nnet.base_nn import SoTaskWrapModule  # noqa: E402
from puresound.nnet.conv_tasnet import TCN, ConvTasNet, GatedTCN  # noqa: E402
from puresound.nnet.dprnn import DPRNN  # noqa: E402
from puresound.nnet.lobe.encoder import ConvEncDec, FreeEncDec  # noqa: E402
from puresound.nnet.lobe.pooling import AttentiveStatisticsPooling  # noqa: E402
from puresound.nnet.lobe.trivial import FiLM, Magnitude, SplitMerge  # noqa: E402
from puresound.nnet.skim import MemLSTM, SegLSTM, SkiM  # noqa: E402
from puresound.nnet.unet import UnetTcn  # noqa: E402
from puresound.nnet.dpcrn import DPCRN, DPRNNblock2D  # noqa: E402
from puresound.nnet.dparn import DPARN  # noqa: E402
from puresound.nnet.lobe.attention import MhaSelfAttenLayer  # noqa: E402

from oracle import describe as D  # noqa: E402
from puresound_b200 import testing as T  # noqa: E402

torch.set_num_threads(8)


def quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def save(name, obj):
    path = os.path.join(HERE, name)
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return scale * (2 * torch.rand(*shape, generator=g) - 1)


def sd_of(m):
    return {k: v.clone() for k, v in m.state_dict().items()}


@torch.no_grad()
def small_cases():
    # ---- lobes ----
    torch.manual_seed(10)
    enc = FreeEncDec(win_length=32, laten_length=24, hop_length=16, output_active=True)
    x = rnd(2, 400, seed=1)
    f = enc(x)
    save("small_free_encdec.pt", {"cfg": D.describe_encoder(enc), "sd": sd_of(enc), "wav": x, "feats": f, "inv": enc.inverse(f)})

    torch.manual_seed(11)
    stft = ConvEncDec(fft_length=64, win_type="hann", win_length=64, hop_length=16, trainable=True, output_format="Complex")
    x = rnd(2, 64 + 16 * 9, seed=2)
    X = stft(x)
    save("small_conv_stft.pt", {"cfg": D.describe_encoder(stft), "sd": sd_of(stft), "wav": x, "spec": X, "inv": stft.inverse(X.clone())})

    cases = {}
    for tag, kw in {
        "gln": dict(causal=False, tcn_norm="gLN", dconv_norm="gGN", emb_dim=0, dilation=2),
        "gln_emb": dict(causal=False, tcn_norm="gLN", dconv_norm="gGN", emb_dim=6, dilation=4),
        "cln_causal": dict(causal=True, tcn_norm="cLN", dconv_norm="cLN", emb_dim=0, dilation=2),
        "bn_causal": dict(causal=True, tcn_norm="bN1d", dconv_norm="bN1d", emb_dim=0, dilation=1),
    }.items():
        torch.manual_seed(12)
        m = T.perturb_(TCN(16, 24, 3, **kw).eval(), seed=3)
        x = rnd(2, 16, 50, seed=4)
        e = rnd(2, 6, seed=5) if kw["emb_dim"] else None
        cases[tag] = {"cfg": D.describe_tcn(m), "sd": sd_of(m), "x": x, "embed": e, "y": m(x, e) if e is not None else m(x)}
    save("small_tcn.pt", cases)

    torch.manual_seed(13)
    m = T.perturb_(ConvTasNet(16, 8, True, tcn_dim=24, per_tcn_stack=3, repeat_tcn=2, tcn_with_embed=[1, 0, 1]).eval(), seed=6)
    x, dv = rnd(2, 16, 70, seed=7), rnd(2, 8, seed=8)
    save("small_conv_tasnet.pt", {"cfg": D.describe_masker(m), "sd": sd_of(m), "x": x, "dvec": dv, "y": m(x, dv)})

    x = rnd(3, 8, 103, seed=9)
    seg, rest = SplitMerge.split(x, 10)
    save("small_splitmerge.pt", {"x": x, "K": 10, "seg": seg, "rest": rest, "merged": SplitMerge.merge(seg, rest)})

    torch.manual_seed(14)
    fm = T.perturb_(FiLM(16, 6, input_norm=True).eval(), seed=10)
    x, c = rnd(4, 16, 10, seed=11), rnd(4, 6, seed=12)
    save("small_film.pt", {"sd": sd_of(fm), "x": x, "cond": c, "y": fm(x, c)})

    cases = {}
    for tag, kw in {
        "overlap_bi": dict(n_blocks=2, seg_size=10, seg_overlap=True, causal=False),
        "nooverlap_causal": dict(n_blocks=2, seg_size=10, seg_overlap=False, causal=True),
        "nooverlap_exact": dict(n_blocks=1, seg_size=10, seg_overlap=False, causal=True),  # T % K == 0 quirk
        "film": dict(n_blocks=3, seg_size=8, seg_overlap=True, causal=True, embed_dim=6, embed_norm=True, block_with_embed=[0, 1, 1]),
    }.items():
        torch.manual_seed(15)
        m = T.perturb_(DPRNN(16, 12, 16, **kw).eval(), seed=13)
        Tn = 60 if tag == "nooverlap_exact" else 57
        x = rnd(2, 16, Tn, seed=14)
        e = rnd(2, 6, seed=15) if kw.get("embed_dim") else None
        cases[tag] = {"cfg": D.describe_masker(m), "sd": sd_of(m), "x": x, "embed": e, "y": m(x, e)}
    torch.manual_seed(16)
    m = T.perturb_(
        DPRNN(16, 12, 16, n_blocks=2, seg_size=10, seg_overlap=False, causal=True, block_with_embed=(False, False), embedding_free_tse=True).eval(),
        seed=16,
    )
    x, e = rnd(2, 16, 57, seed=17), rnd(2, 16, 83, seed=18)
    cases["embedding_free"] = {"cfg": D.describe_masker(m), "sd": sd_of(m), "x": x, "embed": e, "y": m(x, e)}
    save("small_dprnn.pt", cases)

    torch.manual_seed(17)
    pool = T.perturb_(AttentiveStatisticsPooling(16, 8).eval(), seed=19)
    x = rnd(2, 16, 41, seed=20)
    mg = Magnitude(drop_first=False)
    save("small_speaker.pt", {"sd": sd_of(pool), "x": x, "asp": pool(x), "mag": mg(x), "mag_drop": Magnitude(drop_first=True)(x)})

    # ---- wrappers ----
    cases = {}
    torch.manual_seed(18)
    ns = quiet(
        SoTaskWrapModule,
        encoder=FreeEncDec(32, 32, 16),
        masker=ConvTasNet(32, 0, False, tcn_dim=48, per_tcn_stack=4, repeat_tcn=2, tcn_with_embed=[0] * 4),
        mask_constraint="ReLU",
        verbose=False,
    ).eval()
    T.perturb_(ns, seed=21)
    x = T.white(2, 1600, amp=0.1, seed=22)
    cases["ns"] = {"cfg": D.describe(ns), "sd": sd_of(ns), "noisy": x, "enroll": None, "y": ns.inference(x)}

    torch.manual_seed(19)
    tse = quiet(
        SoTaskWrapModule,
        encoder=ConvEncDec(64, "hann", 64, hop_length=16, trainable=True, output_format="Complex"),
        masker=ConvTasNet(64, 12, True, tcn_dim=24, per_tcn_stack=3, repeat_tcn=2, tcn_with_embed=[1, 0, 0]),
        speaker_net=nn.ModuleList(
            [Magnitude(drop_first=False)] + [TCN(32, 24, 3, dilation=2 ** i) for i in range(2)] + [AttentiveStatisticsPooling(32, 8), nn.Conv1d(64, 12, 1, bias=False)]
        ),
        f_type="complex",
        mask_type="complex",
        mask_constraint="linear",
        drop_first_bin=True,
        verbose=False,
    ).eval()
    T.perturb_(tse, seed=23)
    x, e = T.white(2, 64 + 16 * 30, amp=0.1, seed=24), T.white(2, 64 + 16 * 44, amp=0.1, seed=25)
    cases["tse_stft"] = {
        "cfg": D.describe(tse), "sd": sd_of(tse), "noisy": x, "enroll": e, "y": tse.inference(x, e), "dvec": tse.inference_tse_embedding(e),
    }

    torch.manual_seed(20)
    tse2 = quiet(
        SoTaskWrapModule,
        encoder=FreeEncDec(32, 32, 16),
        masker=ConvTasNet(32, 12, True, tcn_dim=24, per_tcn_stack=3, repeat_tcn=2, tcn_with_embed=[1, 0, 0], tcn_norm="bN1d", dconv_norm="bN1d", causal=True),
        speaker_net=nn.ModuleList([TCN(32, 24, 3, dilation=2 ** i) for i in range(2)] + [AttentiveStatisticsPooling(32, 8), nn.Conv1d(64, 12, 1, bias=False)]),
        mask_constraint="ReLU",
        verbose=False,
    ).eval()
    T.perturb_(tse2, seed=26)
    x, e = T.white(2, 1200, amp=0.1, seed=27), T.white(2, 1700, amp=0.1, seed=28)
    cases["tse_free_causal"] = {"cfg": D.describe(tse2), "sd": sd_of(tse2), "noisy": x, "enroll": e, "y": tse2.inference(x, e)}

    torch.manual_seed(21)
    veve = quiet(
        SoTaskWrapModule,
        encoder=FreeEncDec(32, 16, 16, output_active=True),
        masker=DPRNN(16, 12, 16, n_blocks=2, seg_size=10, seg_overlap=False, causal=True, block_with_embed=(False, False), embedding_free_tse=True),
        mask_constraint="ReLU",
        embedding_free_tse=True,
        verbose=False,
    ).eval()
    T.perturb_(veve, seed=29)
    x, e = T.white(2, 1500, amp=0.1, seed=30), T.white(2, 1100, amp=0.1, seed=31)
    cases["veve_dprnn"] = {"cfg": D.describe(veve), "sd": sd_of(veve), "noisy": x, "enroll": e, "y": veve.inference(x, e)}
    save("small_wrappers.pt", cases)


def full_cfgs():
    """SURVEY 8d constructors.  name -> (builder, noisy_len, enroll_len, batch)."""
    def cfg1():
        return quiet(
            SoTaskWrapModule,
            encoder=FreeEncDec(32, 512, 16),
            masker=ConvTasNet(512, 0, False, tcn_kernel=3, tcn_dim=512, repeat_tcn=3, tcn_dilated_basic=2, per_tcn_stack=8, tcn_with_embed=[0] * 8, tcn_norm="gLN", dconv_norm="gGN", causal=False, tcn_layer="normal"),
            mask_constraint="ReLU",
            verbose=False,
        )

    def cfg3():
        return quiet(
            SoTaskWrapModule,
            encoder=FreeEncDec(32, 128, 16, output_active=True),
            masker=DPRNN(128, 128, 128, n_blocks=6, seg_size=100, seg_overlap=True, causal=False),
            mask_constraint="ReLU",
            verbose=False,
        )

    def cfg4():
        return quiet(
            SoTaskWrapModule,
            encoder=ConvEncDec(512, "hann", 512, hop_length=128, trainable=True, output_format="Complex"),
            masker=ConvTasNet(512, 192, True, tcn_dim=256, repeat_tcn=3, per_tcn_stack=8, tcn_with_embed=[1, 0, 0, 0, 0, 0, 0, 0]),
            speaker_net=nn.ModuleList(
                [Magnitude(drop_first=False)] + [TCN(256, 256, 3, dilation=2 ** i) for i in range(5)] + [AttentiveStatisticsPooling(256, 128), nn.Conv1d(512, 192, 1, bias=False)]
            ),
            mask_constraint="linear",
            drop_first_bin=True,
            verbose=False,
        )

    def cfg5():
        return quiet(
            SoTaskWrapModule,
            encoder=FreeEncDec(320, 512, 160),
            masker=ConvTasNet(512, 0, False, tcn_dim=512, per_tcn_stack=8, repeat_tcn=3, tcn_with_embed=[0] * 8, tcn_norm="cLN", dconv_norm="cLN", causal=True),
            mask_constraint="ReLU",
            verbose=False,
        )

    def veve():
        spec = importlib.util.spec_from_file_location("ref_tse_model", os.path.join(REF, "egs/tse/model.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return quiet(mod.init_model, "veve_dprnn_v0_causal", None, None, verbose=False)

    return {
        "cfg1": (cfg1, 64000, None, 2),
        "cfg3": (cfg3, 160000, None, 1),
        "cfg4": (cfg4, 64000, 96000, 2),
        "cfg5_offline": (cfg5, 32000, None, 1),
        "veve_dprnn_v0_causal": (veve, 64000, 96000, 1),
    }


@torch.no_grad()
def full_pins():
    pins = {}
    for name, (build, L, Le, n) in full_cfgs().items():
        torch.manual_seed(0)
        m = build().eval()
        T.perturb_(m, seed=1)
        mix, _ = T.noisy_speech(n, L, seed=1234)
        enr = T.noisy_speech(n, Le, seed=4321)[0] if Le else None
        y = m.inference(mix, enr)
        stride = 997
        pins[name] = {
            "params": sum(p.numel() for p in m.parameters()),
            "state_checksum": T.state_checksum(m.state_dict()),
            "batch": n,
            "length": L,
            "enroll_length": Le,
            "input_seed": 1234,
            "enroll_seed": 4321,
            "stride": stride,
            "out_len": y.shape[-1],
            "out_abs_mean": float(y.abs().mean()),
            "out_clamped_frac": float((y.abs() >= 1).float().mean()),
            "samples": [[float(v) for v in row[::stride]] for row in y],
        }
        print(name, pins[name]["params"], pins[name]["state_checksum"], pins[name]["out_abs_mean"], pins[name]["out_clamped_frac"])
    with open(os.path.join(HERE, "full_size_pins.json"), "w") as fh:
        json.dump(pins, fh)


@torch.no_grad()
def round2_pins():
    """Round-2 additions: the benched shape itself - cfg2 = the cfg1 model at batch 64 (items 0 / 31 / 63 kept) -
    and cfg-1b (the B=128 reading, FreeEncDec(32,16,128) + ConvTasNet(128,...,tcn_dim=512), conv_tasnet.py:239-254)."""
    pins = {}
    stride = 997

    def cfg1b():
        return quiet(
            SoTaskWrapModule,
            encoder=FreeEncDec(32, 128, 16),
            masker=ConvTasNet(128, 0, False, tcn_kernel=3, tcn_dim=512, repeat_tcn=3, tcn_dilated_basic=2, per_tcn_stack=8, tcn_with_embed=[0] * 8, tcn_norm="gLN", dconv_norm="gGN", causal=False, tcn_layer="normal"),
            mask_constraint="ReLU",
            verbose=False,
        )

    for name, build, n, keep in (("cfg1b", cfg1b, 2, [0, 1]), ("cfg2_b64", full_cfgs()["cfg1"][0], 64, [0, 31, 63])):
        torch.manual_seed(0)
        m = build().eval()
        T.perturb_(m, seed=1)
        mix, _ = T.noisy_speech(n, 64000, seed=1234)
        y = m.inference(mix)
        pins[name] = {
            "params": sum(p.numel() for p in m.parameters()),
            "state_checksum": T.state_checksum(m.state_dict()),
            "batch": n, "length": 64000, "enroll_length": None, "input_seed": 1234, "enroll_seed": 4321, "stride": stride,
            "out_len": y.shape[-1], "items": keep,
            "out_abs_mean": float(y[keep].abs().mean()),
            "out_clamped_frac": float((y[keep].abs() >= 1).float().mean()),
            "samples": [[float(v) for v in y[i, ::stride]] for i in keep],
        }
        print(name, pins[name]["params"], pins[name]["state_checksum"], pins[name]["out_abs_mean"], pins[name]["out_clamped_frac"])
    with open(os.path.join(HERE, "round2_pins.json"), "w") as fh:
        json.dump(pins, fh)


def _ref_init_model(name):
    spec = importlib.util.spec_from_file_location("ref_tse_model", os.path.join(REF, "egs/tse/model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return quiet(mod.init_model, name, None, None, verbose=False)


@torch.no_grad()
def skim_cases():
    """SkiM: small variants of the masker and of its two cells, and a full-size pin of the
    reference's demo recipe ``tse_skim_v0_causal`` (egs/tse/model.py:418-463).  Written to their own files so the
    vectors generated earlier stay byte-identical."""
    cases = {}
    for tag, kw in {
        "causal_film": dict(n_blocks=4, seg_size=10, seg_overlap=False, causal=True, embed_dim=6, embed_norm=True,
                            embed_fusion="FiLM", block_with_embed=[1, 1, 0, 1]),
        "bi_overlap": dict(n_blocks=3, seg_size=10, seg_overlap=True, causal=False, embed_dim=0),
        "causal_overlap": dict(n_blocks=2, seg_size=8, seg_overlap=True, causal=True, embed_dim=0),
        "bi_exact": dict(n_blocks=2, seg_size=10, seg_overlap=False, causal=False, embed_dim=0),  # T % K == 0: whole extra segment
    }.items():
        torch.manual_seed(21)
        m = T.perturb_(SkiM(16, 12, 16, **kw).eval(), seed=22)
        Tn = 60 if tag == "bi_exact" else 57
        x = rnd(2, 16, Tn, seed=23)
        e = rnd(2, 6, seed=24) if kw.get("embed_dim") else None
        cases[tag] = {"cfg": D.describe_masker(m), "sd": sd_of(m), "x": x, "embed": e, "y": m(x, e)}
    # the cells on their own, with incoming states (what test/test_streaming.py:10-59 exercises upstream)
    torch.manual_seed(25)
    seg = T.perturb_(SegLSTM(16, 12, causal=True, dropout=0.0).eval(), seed=26)
    x, h, c = rnd(5, 10, 16, seed=27), rnd(1, 5, 12, seed=28), rnd(1, 5, 12, seed=29)
    y, hn, cn = seg(x, h, c)
    cases["seg_cell"] = {"sd": sd_of(seg), "x": x, "h": h, "c": c, "y": y, "hn": hn, "cn": cn}
    for tag, causal in (("mem_cell_causal", True), ("mem_cell_bi", False)):
        torch.manual_seed(30)
        mem = T.perturb_(MemLSTM(12, causal=causal, dropout=0.0).eval(), seed=31)
        Dn = 1 if causal else 2
        h, c = rnd(2, 7, Dn, 12, seed=32), rnd(2, 7, Dn, 12, seed=33)
        ho, co = mem(h, c)
        cases[tag] = {"sd": sd_of(mem), "causal": causal, "h": h, "c": c, "h_out": ho, "c_out": co}
    save("small_skim.pt", cases)

    torch.manual_seed(0)
    m = _ref_init_model("tse_skim_v0_causal").eval()
    T.perturb_(m, seed=1)
    n, L, Le = 1, 64000, 96000
    mix, _ = T.noisy_speech(n, L, seed=1234)
    enr = T.noisy_speech(n, Le, seed=4321)[0]
    y = m.inference(mix, enr)
    stride = 997
    pin = {"params": sum(p.numel() for p in m.parameters()), "state_checksum": T.state_checksum(m.state_dict()), "batch": n,
           "length": L, "enroll_length": Le, "input_seed": 1234, "enroll_seed": 4321, "stride": stride, "out_len": y.shape[-1],
           "out_abs_mean": float(y.abs().mean()), "out_clamped_frac": float((y.abs() >= 1).float().mean()),
           "samples": [[float(v) for v in row[::stride]] for row in y]}
    print("tse_skim_v0_causal", pin["params"], pin["state_checksum"], pin["out_abs_mean"])
    with open(os.path.join(HERE, "skim_pins.json"), "w") as fh:
        json.dump({"tse_skim_v0_causal": pin}, fh)


@torch.no_grad()
def gated_cases():
    """GatedTCN (conv_tasnet.py:93-215): every conditioning mode and norm, an odd channel count, a
    gated Conv-TasNet stack, and a wrapper whose speaker net is built from GatedTCN blocks as in the tse_unet_tcn recipes
    (egs/tse/model.py:226-236).  Own file, so the vectors generated earlier stay byte-identical."""
    cases = {}
    for tag, (cin, hid, kw) in {
        "gln": (16, 24, dict(dilation=2, causal=False, tcn_norm="gLN", emb_dim=0)),
        "gln_concat": (16, 24, dict(dilation=4, causal=False, tcn_norm="gLN", emb_dim=8)),
        "gln_film": (16, 24, dict(dilation=2, causal=False, tcn_norm="gLN", emb_dim=8, use_film=True)),
        "cln_causal_concat": (16, 24, dict(dilation=2, causal=True, tcn_norm="cLN", emb_dim=8)),
        "bn_causal": (16, 24, dict(dilation=1, causal=True, tcn_norm="bN1d", emb_dim=0)),
        "bn_causal_film": (16, 24, dict(dilation=3, causal=True, tcn_norm="bN1d", emb_dim=6, use_film=True)),
        "odd_sizes": (10, 14, dict(dilation=2, causal=False, tcn_norm="gLN", emb_dim=5)),
        "wide_tc": (64, 64, dict(dilation=2, causal=False, tcn_norm="gLN", emb_dim=32)),  # shapes the tcgen05 GEMM takes
    }.items():
        torch.manual_seed(31)
        m = T.perturb_(GatedTCN(cin, hid, 3, **kw).eval(), seed=32)
        x = rnd(2, cin, 300 if tag == "wide_tc" else 50, seed=33)
        e = rnd(2, kw["emb_dim"], seed=34) if kw["emb_dim"] else None
        cases[tag] = {"cfg": D.describe_gated_tcn(m), "sd": sd_of(m), "x": x, "embed": e, "y": m(x, e) if e is not None else m(x)}
    torch.manual_seed(35)
    net = T.perturb_(ConvTasNet(16, 8, True, tcn_layer="gated", tcn_dim=24, per_tcn_stack=3, repeat_tcn=2, tcn_with_embed=[1, 0, 0]).eval(), seed=36)
    x, dv = rnd(2, 16, 70, seed=37), rnd(2, 8, seed=38)
    cases["conv_tasnet_gated"] = {"cfg": D.describe_masker(net), "sd": sd_of(net), "x": x, "dvec": dv, "y": net(x, dv)}
    torch.manual_seed(39)
    wrap = quiet(
        SoTaskWrapModule,
        encoder=ConvEncDec(fft_length=64, win_type="hann", win_length=64, hop_length=16, trainable=True, output_format="Complex"),
        masker=ConvTasNet(64, 12, True, tcn_layer="gated", tcn_dim=24, per_tcn_stack=2, repeat_tcn=2, tcn_with_embed=[1, 0]),
        speaker_net=nn.ModuleList([Magnitude(drop_first=False)] + [GatedTCN(32, 16, 3, dilation=2 ** i, causal=False, tcn_norm="gLN") for i in range(2)]
                                  + [AttentiveStatisticsPooling(32, 8), nn.Conv1d(64, 12, 1, bias=False)]),
        mask_constraint="linear", drop_first_bin=True, verbose=False,
    ).eval()
    T.perturb_(wrap, seed=40)
    noisy, enroll = 0.1 * rnd(2, 64 + 16 * 40, seed=41), 0.1 * rnd(2, 64 + 16 * 55, seed=42)
    cases["wrapper_gated"] = {"cfg": D.describe(wrap), "sd": sd_of(wrap), "noisy": noisy, "enroll": enroll,
                              "y": wrap.inference(noisy, enroll), "dvec": wrap.inference_tse_embedding(enroll)}
    save("small_gated.pt", cases)
    # full size: cfg4 with gated blocks (recipes.baseline_config("cfg4_gated")), pinned like full_size_pins.json
    torch.manual_seed(0)
    m = quiet(
        SoTaskWrapModule,
        encoder=ConvEncDec(512, "hann", 512, hop_length=128, trainable=True, output_format="Complex"),
        masker=ConvTasNet(512, 192, True, tcn_layer="gated", tcn_dim=256, repeat_tcn=3, per_tcn_stack=5, tcn_with_embed=[1, 0, 0, 0, 0]),
        speaker_net=nn.ModuleList([Magnitude(drop_first=False)]
                                  + [GatedTCN(256, 128, 3, dilation=2 ** i, causal=False, tcn_norm="gLN") for i in range(5)]
                                  + [AttentiveStatisticsPooling(256, 128), nn.Conv1d(512, 192, 1, bias=False)]),
        mask_constraint="linear", drop_first_bin=True, verbose=False,
    ).eval()
    T.perturb_(m, seed=1)
    n, L, Le, stride = 2, 64000, 96000, 997
    mix, _ = T.noisy_speech(n, L, seed=1234)
    enr = T.noisy_speech(n, Le, seed=4321)[0]
    y = m.inference(mix, enr)
    pin = {"params": sum(p.numel() for p in m.parameters()), "state_checksum": T.state_checksum(m.state_dict()), "batch": n, "length": L,
           "enroll_length": Le, "input_seed": 1234, "enroll_seed": 4321, "stride": stride, "out_len": y.shape[-1],
           "out_abs_mean": float(y.abs().mean()), "out_clamped_frac": float((y.abs() >= 1).float().mean()),
           "samples": [[float(v) for v in row[::stride]] for row in y]}
    print("cfg4_gated", pin["params"], pin["out_abs_mean"], pin["out_clamped_frac"])
    with open(os.path.join(HERE, "gated_pins.json"), "w") as fh:
        json.dump({"cfg4_gated": pin}, fh)


@torch.no_grad()
def unet_cases():
    """UnetTcn (second half; unet.py:298-556): small variants of the shell, and full-size pins of the
    reference's STFT-domain TSE recipes tse_unet_tcn_v0 / _v0_causal / _v1 (egs/tse/model.py:184-369)."""
    base = dict(input_type="RI", input_dim=32, channels=(1, 4, 8, 8), transpose_t_size=2, kernel_t=(2, 2, 2), kernel_f=(5, 5, 5),
                stride_t=(1, 1, 1), stride_f=(2, 2, 2), dilation_t=(1, 1, 1), dilation_f=(1, 1, 1), delay=(0, 0, 0), tcn_dim=12,
                per_tcn_stack=2, repeat_tcn=2, dropout=0.0)
    cases = {}
    for tag, kw in {
        "gln_gated_concat": dict(embed_dim=6, embed_norm=True, norm_type="gLN", transpose_delay=True, tcn_layer="gated", tcn_with_embed=[1, 0],
                                 tcn_norm="gLN", causal=False),
        "bn_causal_gated_film": dict(embed_dim=6, embed_norm=True, norm_type="bN2d", transpose_delay=True, tcn_layer="gated",
                                     tcn_with_embed=[1, 0], tcn_norm="bN1d", dconv_norm="bN1d", causal=True, tcn_use_film=True),
        "gln_normal_nodelay": dict(embed_dim=0, norm_type="gLN", transpose_delay=False, tcn_layer="normal", tcn_with_embed=[0, 0],
                                   tcn_norm="gLN", dconv_norm="gGN", causal=False),
        "real_lookahead_k3": dict(embed_dim=0, norm_type="bN2d", input_type="Real", input_dim=24, channels=(1, 4, 4, 8), kernel_t=(3, 1, 2),
                                  kernel_f=(3, 5, 1), stride_f=(1, 4, 1), delay=(1, 0, 0), transpose_t_size=1, tcn_layer="normal",
                                  tcn_with_embed=[0, 0], tcn_norm="cLN", dconv_norm="cLN", causal=True),
    }.items():
        torch.manual_seed(51)
        m = T.perturb_(quiet(UnetTcn, **{**base, **kw}).eval(), seed=52)
        cin = kw.get("input_dim", 32)
        x = rnd(2, cin, 37, seed=53)
        e = rnd(2, 6, seed=54) if kw["embed_dim"] else None
        cases[tag] = {"cfg": D.describe_masker(m), "sd": sd_of(m), "x": x, "embed": e, "y": m(x, e) if e is not None else m(x)}
    save("small_unet.pt", cases)
    pins = {}
    for name in ("tse_unet_tcn_v0", "tse_unet_tcn_v0_causal", "tse_unet_tcn_v1"):
        torch.manual_seed(0)
        m = _ref_init_model(name).eval()
        T.perturb_(m, seed=1)
        n, L, Le, stride = 2, 64000, 96000, 997
        mix, _ = T.noisy_speech(n, L, seed=1234)
        enr = T.noisy_speech(n, Le, seed=4321)[0]
        y = m.inference(mix, enr)
        pins[name] = {"params": sum(p.numel() for p in m.parameters()), "state_checksum": T.state_checksum(m.state_dict()), "batch": n,
                      "length": L, "enroll_length": Le, "input_seed": 1234, "enroll_seed": 4321, "stride": stride, "out_len": y.shape[-1],
                      "out_abs_mean": float(y.abs().mean()), "out_clamped_frac": float((y.abs() >= 1).float().mean()),
                      "samples": [[float(v) for v in row[::stride]] for row in y]}
        print(name, pins[name]["params"], pins[name]["out_abs_mean"], pins[name]["out_clamped_frac"])
    with open(os.path.join(HERE, "unet_pins.json"), "w") as fh:
        json.dump(pins, fh)


def _ref_ns_model(name):
    spec = importlib.util.spec_from_file_location("ref_ns_model", os.path.join(REF, "egs/ns/model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return quiet(mod.init_model, name, None, verbose=False)


@torch.no_grad()
def dpcrn_cases():
    """DPCRN (dpcrn.py:11-213): the dual-path block on its own, small variants of the masker, and
    full-size pins of the egs/ns recipes ns_dpcrn_v0 / ns_dpcrn_v0_causal (egs/ns/model.py:38-126)."""
    cases = {}
    torch.manual_seed(61)
    blk = T.perturb_(DPRNNblock2D(16, 12, dropout=0.0).eval(), seed=62)
    x = rnd(2, 16, 9, 11, seed=63)
    cases["block2d"] = {"sd": sd_of(blk), "x": x, "y": blk(x)}
    base = dict(input_type="RI", input_dim=32, channels=(1, 4, 8, 16), transpose_t_size=2, kernel_t=(2, 2, 2), kernel_f=(5, 3, 3),
                stride_t=(1, 1, 1), stride_f=(2, 2, 1), dilation_t=(1, 1, 1), dilation_f=(1, 1, 1), delay=(0, 0, 0), dropout=0.0)
    for tag, kw in {
        "bn_delay_h32": dict(norm_type="bN2d", transpose_delay=True, rnn_hidden=32),     # tensor-core recurrence (H % 32 == 0)
        "bn_causal_h12": dict(norm_type="bN2d", transpose_delay=False, rnn_hidden=12),   # exact-fp32 recurrence
        "gln_h32": dict(norm_type="gLN", transpose_delay=False, rnn_hidden=32),
    }.items():
        torch.manual_seed(64)
        m = T.perturb_(quiet(DPCRN, **{**base, **kw}).eval(), seed=65)
        x = rnd(2, 32, 29, seed=66)
        cases[tag] = {"cfg": D.describe_masker(m), "sd": sd_of(m), "x": x, "y": m(x)}
    save("small_dpcrn.pt", cases)
    pins = {}
    for name in ("ns_dpcrn_v0", "ns_dpcrn_v0_causal"):
        torch.manual_seed(0)
        m = _ref_ns_model(name).eval()
        T.perturb_(m, seed=1)
        n, L, stride = 2, 64000, 997
        mix, _ = T.noisy_speech(n, L, seed=1234)
        y = m.inference(mix)
        pins[name] = {"params": sum(p.numel() for p in m.parameters()), "state_checksum": T.state_checksum(m.state_dict()), "batch": n,
                      "length": L, "input_seed": 1234, "stride": stride, "out_len": y.shape[-1], "out_abs_mean": float(y.abs().mean()),
                      "out_clamped_frac": float((y.abs() >= 1).float().mean()), "samples": [[float(v) for v in row[::stride]] for row in y]}
        print(name, pins[name]["params"], pins[name]["out_abs_mean"], pins[name]["out_clamped_frac"])
    with open(os.path.join(HERE, "dpcrn_pins.json"), "w") as fh:
        json.dump(pins, fh)


def _no_pe(sd):
    """The sin/cos table (5000 x d) is rebuilt by the constructor: keep it out of the fixture."""
    return {k: v for k, v in sd.items() if not k.endswith("pos.pe")}


@torch.no_grad()
def dparn_cases():
    """DPARN (second half; dparn.py:12-246, lobe/attention.py:8-232): the transformer encoder layer
    (with / without positional encoding, causal mask), small variants of the masker, full-size pins of ns_dparn_v0[_causal]."""
    cases = {}
    for tag, (pe, causal) in {"layer_pe": (True, False), "layer_nope_causal": (False, True)}.items():
        torch.manual_seed(71)
        m = T.perturb_(MhaSelfAttenLayer(16, 24, nhead=2, dropout=0.0, improved=False, position_encoding=pe).eval(), seed=72)
        x = rnd(3, 16, 21, seed=73)
        cases[tag] = {"sd": _no_pe(sd_of(m)), "x": x, "pe": pe, "causal": causal, "y": m(x, causal=causal)}
    base = dict(input_type="RI", input_dim=32, channels=(1, 4, 8, 16), transpose_t_size=2, kernel_t=(2, 2, 2), kernel_f=(5, 3, 3),
                stride_t=(1, 1, 1), stride_f=(2, 2, 1), dilation_t=(1, 1, 1), dilation_f=(1, 1, 1), delay=(0, 0, 0), dropout=0.0)
    for tag, kw in {
        "bn_delay_h32_heads4": dict(norm_type="bN2d", transpose_delay=True, rnn_hidden=32, nhead=4),
        "gln_h12_heads1": dict(norm_type="gLN", transpose_delay=False, rnn_hidden=12, nhead=1),
    }.items():
        torch.manual_seed(74)
        m = T.perturb_(quiet(DPARN, **{**base, **kw}).eval(), seed=75)
        x = rnd(2, 32, 29, seed=76)
        cfg = D.describe_masker(m)
        cases[tag] = {"cfg": cfg, "sd": _no_pe(sd_of(m)), "x": x, "y": m(x)}
    save("small_dparn.pt", cases)
    pins = {}
    for name in ("ns_dparn_v0", "ns_dparn_v0_causal"):
        torch.manual_seed(0)
        m = _ref_ns_model(name).eval()
        T.perturb_(m, seed=1)
        n, L, stride = 2, 64000, 997
        mix, _ = T.noisy_speech(n, L, seed=1234)
        y = m.inference(mix)
        pins[name] = {"params": sum(p.numel() for p in m.parameters()), "state_checksum": T.state_checksum(m.state_dict()), "batch": n,
                      "length": L, "input_seed": 1234, "stride": stride, "out_len": y.shape[-1], "out_abs_mean": float(y.abs().mean()),
                      "out_clamped_frac": float((y.abs() >= 1).float().mean()), "samples": [[float(v) for v in row[::stride]] for row in y]}
        print(name, pins[name]["params"], pins[name]["out_abs_mean"], pins[name]["out_clamped_frac"])
    with open(os.path.join(HERE, "dparn_pins.json"), "w") as fh:
        json.dump(pins, fh)


@torch.no_grad()
def sdr_cases():
    """SDR-family scores (loss/sdr.py): every alias of SDRLoss.init_mode that is not source-aggregated,
    per item (reduction=False), and the si_snr function, on estimates at about -5 .. 45 dB and with DC offsets."""
    from puresound.nnet.loss.sdr import SDRLoss, si_snr

    g = torch.Generator().manual_seed(81)
    ref = torch.randn(6, 16000, generator=g) * 0.1 + torch.tensor([0.0, 0.02, -0.05, 0.0, 0.1, 0.0]).unsqueeze(1)
    noise = torch.randn(6, 16000, generator=g) * 0.1
    gains = torch.tensor([1.8, 0.5, 0.1, 0.02, 0.005, 0.3]).unsqueeze(1)
    est = torch.tensor([1.0, 0.7, 1.3, 1.0, 0.9, 1.0]).unsqueeze(1) * ref + gains * noise + 0.01
    out = {"est": est, "ref": ref, "si_snr": si_snr(est, ref, reduction=False), "si_snr_mean": si_snr(est, ref)}
    for mode in ("sisnr", "sdsdr", "sdr", "tsdr"):
        out[mode] = quiet(SDRLoss.init_mode, mode, reduction=False)(est, ref)
        out[mode + "_mean"] = quiet(SDRLoss.init_mode, mode, reduction=True)(est, ref)
    save("small_sdr.pt", out)


@torch.no_grad()
def mel_cases():
    """Mel speaker front-end (second half): FbankEnc (lobe/encoder.py:186-272,459-535), SpecAugment
    (lobe/trivial.py:307-335; random, applied in eval mode too - the global seed set right before each call is stored),
    SingleRNN (lobe/rnn.py:9-52), small wrappers with both speaker nets, and full-size pins of the two recipes that use them
    (tse_skim_v1_causal / tse_skim_v2_causal, egs/tse/model.py:465-549)."""
    from puresound.nnet.lobe.encoder import FbankEnc
    from puresound.nnet.lobe.rnn import SingleRNN
    from puresound.nnet.lobe.trivial import SpecAugment

    cases = {}
    for tag, kw in {"fixed_512": dict(fft_length=512, win_length=512, hop_length=128, trainable=False, n_banks=80),
                    "trainable_128": dict(fft_length=128, win_length=128, hop_length=32, trainable=True, n_banks=20)}.items():
        enc = FbankEnc(output_format="Magnitude", **kw).eval()
        wav = T.noisy_speech(2, 4000, seed=91)[0]
        # the recipe-size tensors (2 x 257 x 512 Fourier kernels) are deterministic: the tests rebuild them from the
        # drop-in constructor (bit-identical, tests/test_host_logic.py) instead of carrying 2 MB here
        cases[tag] = {"cfg": D.describe_encoder(enc), "kw": kw, "sd": sd_of(enc) if tag != "fixed_512" else None, "wav": wav, "mel": enc(wav)}
    for tag, (fm, tm, val, shape) in {"freq": (10, 0, 0.0, (3, 40, 50)), "both": (6, 9, -1.5, (2, 24, 31)), "none": (0, 0, 0.0, (2, 8, 9))}.items():
        x = rnd(*shape, seed=92)
        torch.manual_seed(93)
        cases["specaug_" + tag] = {"freq_mask": fm, "time_mask": tm, "mask_value": val, "x": x, "seed": 93, "y": SpecAugment(fm, tm, val)(x)}
    for tag, bi in (("rnn_bi", True), ("rnn_uni", False)):
        torch.manual_seed(94)
        m = SingleRNN("LSTM", 16, 12, bidirectional=bi, dropout=0.05).eval()
        x = rnd(3, 16, 37, seed=95)
        cases[tag] = {"sd": sd_of(m), "bidirectional": bi, "x": x, "y": m(x)}
    for tag in ("wrapper_mel", "wrapper_rnn"):
        torch.manual_seed(96)
        enc = FreeEncDec(win_length=32, hop_length=16, laten_length=24, output_active=True)
        spk_enc = FbankEnc(fft_length=128, win_length=128, hop_length=32, trainable=False, output_format="Magnitude", n_banks=20) if tag == "wrapper_mel" else None
        masker = SkiM(24, 12, 24, n_blocks=2, seg_size=10, seg_overlap=False, causal=True, embed_dim=6, embed_norm=True,
                      block_with_embed=[1, 1], embed_fusion="FiLM")
        if tag == "wrapper_mel":
            spk = [SpecAugment(freq_mask_length=8, time_mask_length=0, fill_value=0.0)] + [TCN(20, 16, 3, dilation=2 ** i) for i in range(2)] \
                + [AttentiveStatisticsPooling(20, 8), nn.Conv1d(40, 6, 1, bias=False)]
        else:
            spk = [SingleRNN("LSTM", 24, 10, bidirectional=True, dropout=0.05), AttentiveStatisticsPooling(24, 8), nn.Conv1d(48, 6, 1, bias=False)]
        m = quiet(SoTaskWrapModule, encoder=enc, encoder_spk=spk_enc, masker=masker, speaker_net=nn.ModuleList(spk), mask_constraint="ReLU",
                  verbose=False).eval()
        T.perturb_(m, seed=97)
        mix, enr = T.noisy_speech(2, 3200, seed=98)[0], T.noisy_speech(2, 4800, seed=99)[0]
        torch.manual_seed(100)
        y = m.inference(mix, enr)
        torch.manual_seed(100)
        emb = m.inference_tse_embedding(enr)
        cases[tag] = {"cfg": D.describe(m), "sd": sd_of(m), "noisy": mix, "enroll": enr, "seed": 100, "y": y, "emb": emb}
    save("small_mel.pt", cases)

    pins = {}
    for name in ("tse_skim_v1_causal", "tse_skim_v2_causal", "tse_skim_v0_causal_vad"):
        torch.manual_seed(0)
        m = _ref_init_model(name).eval()
        T.perturb_(m, seed=1)
        n, L, Le = 1, 64000, 96000
        mix, _ = T.noisy_speech(n, L, seed=1234)
        enr = T.noisy_speech(n, Le, seed=4321)[0]
        torch.manual_seed(7)
        y = m.inference(mix, enr)
        torch.manual_seed(7)
        emb = m.inference_tse_embedding(enr)
        stride = 997
        pins[name] = {"params": sum(p.numel() for p in m.parameters()), "state_checksum": T.state_checksum(m.state_dict()), "batch": n,
                      "length": L, "enroll_length": Le, "input_seed": 1234, "enroll_seed": 4321, "rng_seed": 7, "stride": stride,
                      "out_len": y.shape[-1], "out_abs_mean": float(y.abs().mean()), "out_clamped_frac": float((y.abs() >= 1).float().mean()),
                      "samples": [[float(v) for v in row[::stride]] for row in y], "embedding": [float(v) for v in emb.flatten()]}
        print(name, pins[name]["params"], pins[name]["state_checksum"], pins[name]["out_abs_mean"])
    with open(os.path.join(HERE, "mel_pins.json"), "w") as fh:
        json.dump(pins, fh)


@torch.no_grad()
def real_input_pins():
    """SURVEY.md 8d inputs (iii) and (i at a = 1.0): the reference's own speech fixture
    (test/test_case/1272-128104-0000_2035-147961-0014.wav, a two-speaker mixture, 16 kHz int16) cropped to 4 s as the
    mixture and tiled to 6 s as the enrollment, and full-scale white noise (output clamp active), through the reference's
    full-size models.  The int16 samples travel inside the fixture so that the GPU box needs neither the reference nor
    its test assets."""
    from scipy.io import wavfile

    sr, pcm = wavfile.read(os.path.join(REF, "test/test_case/1272-128104-0000_2035-147961-0014.wav"))
    assert sr == 16000 and pcm.dtype.name == "int16" and pcm.ndim == 1
    pcm = torch.from_numpy(pcm.copy())
    mix_i16 = pcm[:64000].clone()
    enr_i16 = torch.cat([pcm[20000:], pcm])[:96000].clone()
    mix, enr = (mix_i16.float() / 32768.0)[None], (enr_i16.float() / 32768.0)[None]
    stride = 97
    cfgs = full_cfgs()
    pins = {}
    for tag, name, x, e in (("speech_cfg1", "cfg1", mix, None), ("speech_cfg3", "cfg3", mix, None), ("speech_cfg4", "cfg4", mix, enr),
                            ("speech_veve", "veve_dprnn_v0_causal", mix, enr), ("white_a1_cfg1", "cfg1", T.white(1, 64000, amp=1.0, seed=77), None)):
        torch.manual_seed(0)
        m = cfgs[name][0]().eval()
        T.perturb_(m, seed=1)
        y = m.inference(x, e)
        pins[tag] = {"config": name, "state_checksum": T.state_checksum(m.state_dict()), "stride": stride, "out_len": y.shape[-1],
                     "out_abs_mean": float(y.abs().mean()), "out_clamped_frac": float((y.abs() >= 1).float().mean()),
                     "samples": [float(v) for v in y[0, ::stride]]}
        print(tag, pins[tag]["out_abs_mean"], pins[tag]["out_clamped_frac"])
    save("real_input_pins.pt", {"mix_i16": mix_i16, "enroll_i16": enr_i16, "white_seed": 77, "pins": pins})


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "small"):
        small_cases()
    if which in ("all", "full"):
        full_pins()
    if which in ("all", "round2"):
        round2_pins()
    if which in ("all", "skim"):
        skim_cases()
    if which in ("all", "real"):
        real_input_pins()
    if which in ("all", "gated"):
        gated_cases()
    if which in ("all", "unet"):
        unet_cases()
    if which in ("all", "dpcrn"):
        dpcrn_cases()
    if which in ("all", "dparn"):
        dparn_cases()
    if which in ("all", "sdr"):
        sdr_cases()
    if which in ("all", "mel"):
        mel_cases()
