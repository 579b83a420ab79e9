"""Drop-in modules on the B200 against the REFERENCE's own outputs (tests/golden/*.pt) on identical weights and
inputs, called through the reference's API ([N,C,T] tensors, inference(noisy, enroll))."""
import pytest
import torch

import _build
from puresound_b200.nnet.lobe.pooling import AttentiveStatisticsPooling
from puresound_b200.nnet.lobe.trivial import FiLM, Magnitude, SplitMerge

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-5  # exact-fp32 kernels: summation-order noise only


def close(a, b, tol=TOL):
    a, b = a.cpu(), b.cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), err


def cu(t):
    return None if t is None else t.to(DEV)


def test_free_encdec(golden):
    g = golden("small_free_encdec.pt")
    m = _build.encoder(g["cfg"]).to(DEV)
    m.load_state_dict(g["sd"])
    f = m(cu(g["wav"]))
    close(f, g["feats"])
    close(m.inverse(f), g["inv"])
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 8, device=DEV))


def test_conv_stft(golden):
    g = golden("small_conv_stft.pt")
    m = _build.encoder(g["cfg"]).to(DEV)
    m.load_state_dict(g["sd"])
    X = m(cu(g["wav"]))
    close(X, g["spec"])
    y = m.inverse(X)
    close(y, g["inv"], 1e-4)  # window-sumsquare division amplifies rounding at the edges
    assert y[:, 0].abs().max().item() == 0.0


def test_tcn_variants(golden):
    for tag, g in golden("small_tcn.pt").items():
        m = _build.tcn(g["cfg"]).to(DEV).eval()
        m.load_state_dict(g["sd"])
        y = m(cu(g["x"]), cu(g["embed"])) if g["embed"] is not None else m(cu(g["x"]))
        close(y, g["y"])


@pytest.mark.parametrize("backend", ["auto", "simt"])
def test_gated_tcn(golden, backend):
    """GatedTCN on the engine against the reference's outputs: concat / FiLM / no conditioning,
    gLN / cLN / bN1d, causal trim, channel counts that are not multiples of 4 (scalar `ps_gated` path), shapes the tcgen05
    GEMM takes (`wide_tc`, auto back end), a gated Conv-TasNet stack and a wrapper whose speaker net is GatedTCN blocks."""
    from puresound_b200 import ops

    ops.force_gemm_backend = ops.GEMM_SIMT if backend == "simt" else None
    try:
        gs = golden("small_gated.pt")
        for tag, g in gs.items():
            if tag in ("conv_tasnet_gated", "wrapper_gated"):
                continue
            m = _build.gated_tcn(g["cfg"]).to(DEV).eval()
            m.load_state_dict(g["sd"])
            y = m(cu(g["x"]), cu(g["embed"])) if g["embed"] is not None else m(cu(g["x"]))
            close(y, g["y"], 5e-5 if tag == "wide_tc" else TOL)
        g = gs["conv_tasnet_gated"]
        m = _build.masker(g["cfg"]).to(DEV).eval()
        m.load_state_dict(g["sd"])
        close(m(cu(g["x"]), cu(g["dvec"])), g["y"])
        g = gs["wrapper_gated"]
        m = _build.wrapper(g["cfg"], g["sd"]).to(DEV)
        close(m.inference(cu(g["noisy"]), cu(g["enroll"])), g["y"], 1e-4)
        close(m.inference_tse_embedding(cu(g["enroll"])), g["dvec"], 5e-5)
        close(m.inference(g["noisy"], g["enroll"]), g["y"], 1e-4)  # third call: CUDA-graph replay, host buffers
    finally:
        ops.force_gemm_backend = None


@pytest.mark.parametrize("backend", ["auto", "simt"])
def test_unet_tcn(golden, backend):
    """UnetTcn shell on the engine against the reference's outputs: 2-D convs as framed GEMMs over tap
    buffers, transposed convs per output phase, gLN / bN2d, time trims, gated / normal bottlenecks."""
    from puresound_b200 import ops

    ops.force_gemm_backend = ops.GEMM_SIMT if backend == "simt" else None
    try:
        for tag, g in golden("small_unet.pt").items():
            m = _build.masker(g["cfg"]).to(DEV).eval()
            m.load_state_dict(g["sd"])
            y = m(cu(g["x"]), cu(g["embed"])) if g["embed"] is not None else m(cu(g["x"]))
            close(y, g["y"], 5e-5)
    finally:
        ops.force_gemm_backend = None


def test_dpcrn(golden):
    """DPCRN on the engine against the reference's outputs: the 2-D dual-path block (bidirectional
    LSTM over frequency, uni-directional over time, in place on [N, T, F, C]) and the masker inside bN2d / gLN shells."""
    from puresound_b200.nnet.dpcrn import DPRNNblock2D

    gs = golden("small_dpcrn.pt")
    blk = DPRNNblock2D(16, 12).to(DEV).eval()
    blk.load_state_dict(gs["block2d"]["sd"])
    close(blk(cu(gs["block2d"]["x"])), gs["block2d"]["y"], 5e-5)
    for tag, g in gs.items():
        if tag == "block2d":
            continue
        m = _build.masker(g["cfg"]).to(DEV).eval()
        m.load_state_dict(g["sd"])
        close(m(cu(g["x"])), g["y"], 1e-4)


def test_dparn(golden):
    """DPARN on the engine against the reference's outputs: the transformer encoder layer (attention
    kernel, positional encoding as a GEMM residual table, causal mask) and the masker inside bN2d / gLN shells."""
    from puresound_b200.nnet.dparn import MhaSelfAttenLayer

    gs = golden("small_dparn.pt")
    for tag in ("layer_pe", "layer_nope_causal"):
        g = gs[tag]
        m = MhaSelfAttenLayer(16, 24, nhead=2, position_encoding=g["pe"]).to(DEV).eval()
        assert not m.load_state_dict(g["sd"], strict=False).unexpected_keys
        close(m(cu(g["x"]), causal=g["causal"]), g["y"], 5e-5)
    for tag, g in gs.items():
        if tag.startswith("layer"):
            continue
        m = _build.masker(g["cfg"]).to(DEV).eval()
        missing = m.load_state_dict(g["sd"], strict=False)
        assert not missing.unexpected_keys and all(k.endswith("pos.pe") for k in missing.missing_keys)
        close(m(cu(g["x"])), g["y"], 1e-4)


def test_attention_kernel():
    """ps_attention against torch's scaled_dot_product_attention: head dims 4..64, ragged lengths, causal."""
    from puresound_b200 import ops

    g = torch.Generator().manual_seed(3)
    # the last two cases exceed the whole-sequence kernel's shared-memory tile and take the per-(sequence, head) kernel
    for B, L, E, H, causal in [(5, 64, 128, 8, False), (3, 37, 64, 2, True), (2, 130, 48, 12, False), (4, 9, 64, 1, True),
                               (1, 200, 384, 12, False), (2, 150, 384, 6, True)]:
        qkv = (2 * torch.rand(B, L, 3 * E, generator=g) - 1).to(DEV)
        out = ops.attention(qkv, H, causal)
        q, k, v = [t.view(B, L, H, E // H).transpose(1, 2).double() for t in qkv.split(E, dim=-1)]
        ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal).transpose(1, 2).reshape(B, L, E)
        assert (out.double() - ref).abs().max().item() <= 2e-6


def test_sdr_scores_on_device(golden):
    """`ps_sdr` (one pass, fp64 moments) behind the reference's SDRLoss / si_snr API against the reference's outputs: every
    non-aggregated alias, per item and reduced, DC offsets, -5 .. 45 dB; tolerance 1e-3 dB."""
    from puresound_b200.nnet.loss.sdr import SDRLoss, align_waveform, si_snr

    g = golden("small_sdr.pt")
    est, ref = cu(g["est"]), cu(g["ref"])
    for mode in ("sisnr", "sdsdr", "sdr", "tsdr"):
        got = SDRLoss.init_mode(mode, reduction=False)(est, ref).cpu()
        assert got.shape == g[mode].shape and (got - g[mode]).abs().max().item() <= 1e-3, mode
        assert abs(float(SDRLoss.init_mode(mode)(est, ref)) - float(g[mode + "_mean"])) <= 1e-3
    assert (si_snr(est, ref, reduction=False).cpu().view(-1) - g["si_snr"].view(-1)).abs().max().item() <= 1e-3
    assert abs(float(si_snr(est, ref)) - float(g["si_snr_mean"])) <= 1e-3
    with pytest.raises(NameError):
        SDRLoss.init_mode("nope")
    with pytest.raises(NotImplementedError):
        SDRLoss.init_mode("sasdr")
    a, b = align_waveform(est, ref[:, :15000])           # shorter reference: left-padded (base_nn.py:404-408)
    assert b.shape == a.shape and bool((b[:, :1000] == 0).all()) and torch.equal(b[:, 1000:], ref[:, :15000])
    a, b = align_waveform(est[:, :12000], ref)
    assert torch.equal(b, ref[:, :12000])


def test_conv_tasnet(golden):
    g = golden("small_conv_tasnet.pt")
    m = _build.masker(g["cfg"]).to(DEV).eval()
    m.load_state_dict(g["sd"])
    close(m(cu(g["x"]), cu(g["dvec"])), g["y"])


def test_splitmerge_film_speaker(golden):
    g = golden("small_splitmerge.pt")
    seg, rest = SplitMerge.split(cu(g["x"]), g["K"])
    assert rest == g["rest"] and torch.equal(seg.cpu(), g["seg"])
    assert torch.equal(SplitMerge.merge(seg, rest).cpu(), g["merged"])
    g = golden("small_film.pt")
    fm = FiLM(16, 6).to(DEV)
    fm.load_state_dict(g["sd"])
    close(fm(cu(g["x"]), cu(g["cond"])), g["y"])
    g = golden("small_speaker.pt")
    pool = AttentiveStatisticsPooling(16, 8).to(DEV).eval()
    pool.load_state_dict(g["sd"])
    close(pool(cu(g["x"])), g["asp"])
    close(Magnitude(drop_first=False)(cu(g["x"])), g["mag"])
    close(Magnitude(drop_first=True)(cu(g["x"])), g["mag_drop"])
    with pytest.raises(NotImplementedError):
        pool.train()(cu(g["x"]))


def test_dprnn_variants(golden):
    for tag, g in golden("small_dprnn.pt").items():
        c = dict(g["cfg"])
        c["output_size"] = g["sd"]["output_fc.1.weight"].shape[0]
        m = _build.masker(c).to(DEV).eval()
        m.load_state_dict(g["sd"])
        y = m(cu(g["x"]), cu(g["embed"])) if g["embed"] is not None else m(cu(g["x"]))
        close(y, g["y"], 5e-5)


def test_skim_variants(golden):
    """SkiM and its cells on the engine against the reference's outputs."""
    from puresound_b200.nnet.skim import MemLSTM, SegLSTM

    gs = golden("small_skim.pt")
    for tag in ("causal_film", "bi_overlap", "causal_overlap", "bi_exact"):
        g = gs[tag]
        c = dict(g["cfg"])
        c["output_size"] = g["sd"]["output_fc.1.weight"].shape[0]
        m = _build.masker(c).to(DEV).eval()
        m.load_state_dict(g["sd"])
        y = m(cu(g["x"]), cu(g["embed"])) if g["embed"] is not None else m(cu(g["x"]))
        close(y, g["y"], 5e-5)
    g = gs["seg_cell"]
    seg = SegLSTM(16, 12, causal=True).to(DEV).eval()
    seg.load_state_dict(g["sd"])
    y, hn, cn = seg(cu(g["x"]), cu(g["h"]), cu(g["c"]))
    close(y, g["y"], 2e-5)
    close(hn, g["hn"], 2e-5)
    close(cn, g["cn"], 2e-5)
    for tag in ("mem_cell_causal", "mem_cell_bi"):
        g = gs[tag]
        mem = MemLSTM(12, causal=g["causal"]).to(DEV).eval()
        mem.load_state_dict(g["sd"])
        ho, co = mem(cu(g["h"]), cu(g["c"]))
        close(ho, g["h_out"], 2e-5)
        close(co, g["c_out"], 2e-5)


def test_mel_front_end(golden):
    """Mel speaker front-end (second half) on the engine against the reference's outputs: FbankEnc (both
    GEMM back ends for the analysis), SpecAugment under the recorded seed in both layouts, SingleRNN, and the wrappers."""
    from puresound_b200.nnet.lobe.encoder import FbankEnc
    from puresound_b200.nnet.lobe.rnn import SingleRNN
    from puresound_b200.nnet.lobe.trivial import SpecAugment

    gs = golden("small_mel.pt")
    for tag in ("fixed_512", "trainable_128"):
        g = gs[tag]
        enc = FbankEnc(output_format="Magnitude", **g["kw"]).to(DEV).eval()
        if g["sd"] is not None:
            enc.load_state_dict(g["sd"])
        close(enc(cu(g["wav"])), g["mel"], 2e-5)  # exact-fp32 analysis
        mel_tc = enc.encode_cl(cu(g["wav"]), exact=False).transpose(1, 2)  # tcgen05 analysis where the shape allows
        close(mel_tc, g["mel"], 1e-4)
    for tag in ("specaug_freq", "specaug_both", "specaug_none"):
        g = gs[tag]
        aug = SpecAugment(g["freq_mask"], g["time_mask"], g["mask_value"])
        torch.manual_seed(g["seed"])
        y = aug(cu(g["x"]))
        assert torch.equal(y.cpu(), g["y"])
        torch.manual_seed(g["seed"])
        ycl = aug.forward_cl(cu(g["x"]).transpose(1, 2).contiguous()).transpose(1, 2)
        assert torch.equal(ycl.cpu(), g["y"])
    for tag in ("rnn_bi", "rnn_uni"):
        g = gs[tag]
        m = SingleRNN("LSTM", 16, 12, bidirectional=g["bidirectional"]).to(DEV).eval()
        m.load_state_dict(g["sd"])
        close(m(cu(g["x"])), g["y"], 2e-5)
    for tag in ("wrapper_mel", "wrapper_rnn"):
        g = gs[tag]
        m = _build.wrapper(g["cfg"], g["sd"]).to(DEV)
        for rep in range(3):  # eager, graph capture, graph replay: every call draws its band like the reference would
            torch.manual_seed(g["seed"])
            close(m.inference(cu(g["noisy"]), cu(g["enroll"])), g["y"], 1e-4)
        torch.manual_seed(g["seed"])
        close(m.inference_tse_embedding(cu(g["enroll"])), g["emb"], 5e-5)
    # a different seed moves the band: the replayed graph must follow the host's draw, not the captured one
    g = gs["wrapper_mel"]
    m = _build.wrapper(g["cfg"], g["sd"]).to(DEV)
    from oracle import separator_ref as R
    for seed in (1, 2, 3, 4):
        torch.manual_seed(seed)
        y = m.inference(cu(g["noisy"]), cu(g["enroll"]))
        torch.manual_seed(seed)
        close(y, R.inference(g["sd"], g["cfg"], g["noisy"], g["enroll"]), 1e-4)


def test_wrappers_inference(golden):
    for tag, g in golden("small_wrappers.pt").items():
        m = _build.wrapper(g["cfg"], g["sd"]).to(DEV)
        y = m.inference(cu(g["noisy"]), cu(g["enroll"]))
        close(y, g["y"], 1e-4)
        # host-buffer API: same numbers, result back on the host
        yh = m.inference(g["noisy"], g["enroll"])
        assert not yh.is_cuda
        close(yh, g["y"], 1e-4)
        if "dvec" in g:
            close(m.inference_tse_embedding(cu(g["enroll"])), g["dvec"], 5e-5)


@pytest.mark.parametrize("name,lookahead,receptive", [
    ("td_tse_conv_tasnet_v0_causal", "16", "24496"),   # 1530 frames x 16 + 16
    ("veve_dprnn_v0_causal", "16", "infinite"),        # egs/tse/model.py:609-613
    # the widening rows: numbers printed by the reference's own _verbose() on these recipes
    ("tse_skim_v0_causal", "16", "infinite"),          # egs/tse/model.py:418-423
    ("tse_unet_tcn_v0_causal", "1152", "24960"),       # egs/tse/model.py:245-250
    ("ns_dpcrn_v0_causal", "384", "infinite"),         # egs/ns/model.py:38-43
    ("ns_dpcrn_v0", "1024", "infinite"),               # egs/ns/model.py:84-89 (semi-causal: transpose_delay)
    ("ns_dparn_v0_causal", "384", "infinite"),         # egs/ns/model.py:128-133
    ("tse_skim_v1_causal", "16", "infinite"),          # egs/tse/model.py:465-470
    ("tse_skim_v2_causal", "16", "infinite"),          # egs/tse/model.py:509-514 (mel speaker front-end, SpecAugment)
    ("tse_skim_v0_causal_vad", "16", "infinite"),      # egs/tse/model.py:560-565
])
def test_verbose_probe_known_answers(name, lookahead, receptive, capsys):
    """The reference's `_verbose()` probe (base_nn.py:740-777) feeds +inf into half of a 10 s signal and reads look-ahead /
    receptive field off the first / last NaN of the output: every kernel on the path (tcgen05 GEMMs, depthwise conv,
    norms, LSTM, overlap-add, the CUDA-graph replay) must propagate Inf/NaN exactly like ATen for the printed numbers to
    match the recipe docstrings."""
    from puresound_b200 import recipes

    torch.manual_seed(0)
    m = recipes.init_model(name, verbose=False).to("cuda")
    m._verbose()
    out = capsys.readouterr().out
    assert f"Lookahead(samples): {lookahead}" in out, out
    assert f"Receptive Fields(samples): {receptive}" in out, out
