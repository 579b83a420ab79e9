"""Streaming == offline, the equivalence the reference pins for its own streaming model (test/test_streaming.py:62-116).
Oracle: the offline causal forward (oracle.inference) of the same weights on the whole signal."""
import pytest
import torch

from oracle import describe as D
from oracle import separator_ref as R
from puresound_b200 import testing
from puresound_b200.nnet.base_nn import SoTaskWrapModule
from puresound_b200.nnet.lobe.encoder import FreeEncDec
from puresound_b200.streaming.conv_tasnet_inference import StreamingConvTasNet, StreamingSeparator

pytestmark = pytest.mark.gpu


def build(norm, win, hop, nf, hid, X, Rp, seed):
    torch.manual_seed(seed)
    m = SoTaskWrapModule(
        FreeEncDec(win, nf, hop),
        StreamingConvTasNet(nf, 0, False, tcn_dim=hid, per_tcn_stack=X, repeat_tcn=Rp, tcn_with_embed=[0] * X, tcn_norm=norm, dconv_norm=norm, causal=True),
        mask_constraint="ReLU", verbose=False).eval()
    testing.perturb_(m, seed=seed + 1)
    return m


def run_stream(m, wav, use_graph, use_hop_kernel=True):
    S, L = wav.shape
    sep = StreamingSeparator(m, use_graph=use_graph, use_hop_kernel=use_hop_kernel)
    sep.init_status(S)
    assert (sep._hop is not None) == use_hop_kernel
    hop, win = sep.hop, sep.win
    outs = []
    for j in range(L // hop):
        outs.append(sep.step_wave(wav[:, j * hop:(j + 1) * hop].cuda()))
    y = torch.cat(outs, dim=1).cpu()
    return y[:, (win // hop - 1) * hop:]  # drop the priming hops


@pytest.mark.parametrize("norm,win,hop", [("cLN", 32, 16), ("bN1d", 32, 16), ("cLN", 48, 16)])
@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("hop_kernel", [True, False])
def test_streaming_equals_offline_small(norm, win, hop, use_graph, hop_kernel):
    """hop_kernel=True: the persistent cooperative kernel (ps_stream_hop, one launch per hop); False: the kernel chain."""
    m = build(norm, win, hop, 24, 40, 4, 2, seed=3)
    wav = testing.white(3, hop * 90, amp=0.1, seed=5)
    ref = R.inference(m.state_dict(), D.describe(m), wav)
    y = run_stream(m.cuda(), wav, use_graph, hop_kernel)
    n = y.shape[1]
    assert n >= ref.shape[1] - win
    err = (y - ref[:, :n]).abs().max().item()
    assert err <= 2e-5, err
    # and the offline engine path of the same module agrees too
    off = m.inference(wav)
    assert (off - ref).abs().max().item() <= 2e-5


def test_streaming_cfg5_full_width():
    """cfg-5 architecture (N=512, H=512, X=8, R=3, cLN, 10 ms hop): 8 streams x 40 hops against the offline oracle."""
    from puresound_b200 import recipes

    torch.manual_seed(0)
    ref_model = recipes.baseline_config("cfg5").eval()
    testing.perturb_(ref_model, seed=1)
    m = build("cLN", 320, 160, 512, 512, 8, 3, seed=0)
    m.load_state_dict(ref_model.state_dict())
    wav = testing.noisy_speech(8, 160 * 41, seed=11)[0]
    ref = R.inference(m.state_dict(), D.describe(m), wav)
    y = run_stream(m.cuda(), wav, True)
    n = y.shape[1]
    err = (y - ref[:, :n]).abs().max().item()
    print(f"cfg5 streaming vs offline oracle: max|err|={err:.3e} over {n} samples x 8 streams")
    assert err <= 1e-3


def test_streaming_many_streams_tensor_core_path():
    """From TC_MIN_STREAMS concurrent streams on, the per-hop 1x1 convs run on the CTA-pair tcgen05 kernel (3xBF16):
    64 streams x 24 hops of the cfg-5 width against the offline oracle, north_star tolerance (1e-3)."""
    m = build("cLN", 320, 160, 512, 512, 4, 1, seed=2)
    assert StreamingConvTasNet.TC_MIN_STREAMS <= 64
    wav = testing.noisy_speech(64, 160 * 25, seed=12)[0]
    ref = R.inference(m.state_dict(), D.describe(m), wav)
    y = run_stream(m.cuda(), wav, True)
    n = y.shape[1]
    err = (y - ref[:, :n]).abs().max().item()
    print(f"64-stream tensor-core hop vs offline oracle: max|err|={err:.3e}")
    assert err <= 1e-3


@pytest.mark.parametrize("norm,use_graph", [("cLN", True), ("bN1d", True), ("bN1d", False)])
def test_streaming_with_speaker_embedding_equals_offline(norm, use_graph):
    """Speaker-conditioned streaming (reference signature step_frame(x, embed), skim_inference.py:177; embed concat
    conv_tasnet.py:78-83): a small causal TSE model - conditioned block 0 of each repeat, embed_norm, TCN + ASP speaker
    net - streamed with one enrollment per stream equals the offline oracle run with the same enrollments; then the
    embeddings are swapped between streams on the live separator (set_embedding) and the outputs follow."""
    from puresound_b200.nnet.conv_tasnet import TCN, ConvTasNet
    from puresound_b200.nnet.lobe.pooling import AttentiveStatisticsPooling

    torch.manual_seed(7)
    nf, hid, E = 24, 40, 12
    m = SoTaskWrapModule(
        FreeEncDec(32, nf, 16),
        ConvTasNet(nf, E, True, tcn_dim=hid, per_tcn_stack=3, repeat_tcn=2, tcn_with_embed=[1, 0, 0], tcn_norm=norm, dconv_norm=norm, causal=True),
        speaker_net=torch.nn.ModuleList([TCN(nf, 16, 3, dilation=2 ** i) for i in range(2)] + [AttentiveStatisticsPooling(nf, 8), torch.nn.Conv1d(2 * nf, E, 1, bias=False)]),
        mask_constraint="ReLU", verbose=False).eval()
    testing.perturb_(m, seed=8)
    S, hop, win = 3, 16, 32
    wav = testing.white(S, hop * 70, amp=0.1, seed=9)
    enr = testing.white(S, 1600, amp=0.1, seed=10)
    sd, cfg = {k: v.clone() for k, v in m.state_dict().items()}, D.describe(m)
    ref = R.inference(sd, cfg, wav, enr)
    m = m.cuda()
    sep = StreamingSeparator(m, use_graph=use_graph)
    sep.init_status(S, enroll=enr)

    def stream(sep):
        outs = [sep.step_wave(wav[:, j * hop:(j + 1) * hop].cuda()) for j in range(wav.shape[1] // hop)]
        return torch.cat(outs, dim=1).cpu()[:, (win // hop - 1) * hop:]

    y = stream(sep)
    n = y.shape[1]
    err = (y - ref[:, :n]).abs().max().item()
    assert err <= 2e-5, err
    assert (m.inference(wav, enr) - ref).abs().max().item() <= 2e-5  # offline engine path of the same model
    # ready-made embeddings, rolled by one stream: every stream now follows its neighbour's speaker
    dvec = m.inference_tse_embedding(enr.cuda()).squeeze(-1)
    ref_rolled = R.inference(sd, cfg, wav, enr.roll(1, 0))
    sep.init_status(S, embed=dvec.roll(1, 0))
    y2 = stream(sep)
    assert (y2 - ref_rolled[:, :n]).abs().max().item() <= 2e-5
    assert (y2 - y).abs().max().item() > 1e-5  # the conditioning matters (white-noise enrollments give close embeddings)
    with pytest.raises(ValueError):
        sep.init_status(S)  # a conditioned model without enrollment


def test_streaming_td_tse_conv_tasnet_v0_causal_recipe():
    """The one causal Conv-TasNet recipe the reference ships (egs/tse/model.py:142-182, bN1d norms, dvec 192 into block 0 of
    each repeat) built by recipes.init_model as an OFFLINE model and streamed through its streaming view: 2 streams x 0.5 s
    with a 1.5 s enrollment each against the offline oracle."""
    from puresound_b200 import recipes

    torch.manual_seed(0)
    m = recipes.init_model("td_tse_conv_tasnet_v0_causal", verbose=False).eval()
    testing.perturb_(m, seed=1)
    wav = testing.noisy_speech(2, 16 * 500, seed=21)[0]
    enr = testing.noisy_speech(2, 24000, seed=22)[0]
    ref = R.inference(m.state_dict(), D.describe(m), wav, enr)
    m = m.cuda()
    sep = StreamingSeparator(m, use_graph=True)
    sep.init_status(2, enroll=enr)
    outs = [sep.step_wave(wav[:, j * 16:(j + 1) * 16].cuda()) for j in range(500)]
    y = torch.cat(outs, dim=1).cpu()[:, 16:]
    n = y.shape[1]
    err = (y - ref[:, :n]).abs().max().item()
    print(f"td_tse_conv_tasnet_v0_causal streaming vs offline oracle: max|err|={err:.3e} over {n} samples x 2 streams")
    assert err <= 1e-3


@pytest.mark.parametrize("n_streams", [1, 3])
def test_streaming_skim_equals_offline(n_streams):
    """The reference's own streaming test (test/test_streaming.py:62-116): StreamingSkiM offline forward == step_chunk ==
    step_frame (mean abs error < 1e-7 upstream, on one stream).  Here also against the oracle and for several streams.

    Several streams are INDEPENDENT streams, i.e. each equals the offline model run on that item alone.  The reference's
    batched offline forward is not that: its causal MemLSTM shifts the memories by one along the flattened [N*S] axis
    (skim.py:103-110), so item n's first segment starts from item n-1's last memory.  The offline engine reproduces the
    shift bit for bit (checked below against the batched oracle); the streams are checked against per-item oracle runs."""
    from puresound_b200.streaming.skim_inference import StreamingSkiM

    torch.manual_seed(4)
    model = StreamingSkiM(5, 20, 5, seg_size=10, seg_overlap=False, causal=True, n_blocks=4, embed_dim=10, embed_norm=True,
                          embed_fusion="FiLM", block_with_embed=[1, 1, 1, 1]).eval()
    testing.perturb_(model, seed=5)
    S, T = n_streams, 1000
    x, d = torch.rand(S, 5, T), torch.rand(S, 10)
    desc = D.describe_masker(model)
    ref_batched = R.skim(model.state_dict(), "", x, d, desc)
    ref = torch.cat([R.skim(model.state_dict(), "", x[s:s + 1], d[s:s + 1], desc) for s in range(S)], 0)
    model = model.cuda()
    assert (model(x.cuda(), d.cuda()).cpu() - ref_batched).abs().max().item() <= 2e-5
    y1 = torch.cat([model(x[s:s + 1].cuda(), d[s:s + 1].cuda()) for s in range(S)], 0).cpu()
    assert (y1 - ref).abs().max().item() <= 2e-5
    if S > 1:
        assert (ref_batched[1:, :, :10] - ref[1:, :, :10]).abs().max().item() > 1e-3  # the upstream cross-item leak is real
    # chunk by chunk with explicit state passing (skim_inference.py:43-139)
    xt = x.cuda().transpose(1, 2).contiguous()  # [S, T, C]
    sh = mh = sc = mc = None
    outs = []
    for k in range(T // 10):
        o, sh, mh, sc, mc = model.step_chunk(xt[:, k * 10:(k + 1) * 10, :], sh, mh, sc, mc, d.cuda())
        outs.append(o)
    y2 = torch.cat(outs, dim=-1).cpu()
    assert (y2 - ref).abs().max().item() <= 2e-5 and (y2 - y1).abs().mean().item() < 1e-6
    # frame by frame with the managed state (skim_inference.py:142-252)
    model.init_status(S)
    outs = [model.step_frame(xt[:, t:t + 1, :], d.cuda()) for t in range(T)]
    y3 = torch.cat(outs, dim=-1).cpu()
    assert (y3 - ref).abs().max().item() <= 2e-5 and (y3 - y1).abs().mean().item() < 1e-6


def test_streaming_guards():
    with pytest.raises(AssertionError):
        StreamingConvTasNet(16, 0, tcn_dim=8, per_tcn_stack=1, repeat_tcn=1, tcn_with_embed=[0], causal=False)
    with pytest.raises(AssertionError):
        StreamingConvTasNet(16, 0, tcn_dim=8, per_tcn_stack=1, repeat_tcn=1, tcn_with_embed=[0], causal=True, tcn_norm="gLN", dconv_norm="cLN")
    m = build("cLN", 32, 16, 16, 16, 1, 1, seed=1).cuda()
    sep = StreamingSeparator(m)
    with pytest.raises(RuntimeError):
        sep.step_wave(torch.zeros(1, 16))
    sep.init_status(2)
    with pytest.raises(ValueError):
        sep.step_wave(torch.zeros(1, 16))


@pytest.mark.parametrize("S", [1, 5, 40, 200, 256])
def test_hop_kernel_stream_counts_and_launches(S):
    """ps_stream_hop at stream counts that exercise every tile shape of its GEMM phases (1 x 4 ... 16 x 64 channel slabs, ragged
    last row group), full cfg-5 width, against the kernel chain on the same inputs (both exact fp32: 2e-5) - and ONE launch per
    hop instead of ~125."""
    from puresound_b200 import ops

    m = build("cLN", 320, 160, 512, 512, 3, 1, seed=4).cuda()
    wav = testing.white(S, 160 * 8, amp=0.1, seed=6)
    outs = {}
    for hopk in (True, False):
        sep = StreamingSeparator(m, use_graph=False, use_hop_kernel=hopk)
        sep.init_status(S)
        ys, n0 = [], None
        for j in range(8):
            if j == 7:
                n0 = ops.launch_count
            ys.append(sep.step_wave(wav[:, j * 160:(j + 1) * 160].cuda()))
        outs[hopk] = (torch.cat(ys, dim=1).cpu(), ops.launch_count - n0)
    assert (outs[True][0] - outs[False][0]).abs().max().item() <= (2e-5 if S < 32 else 2e-4)  # (tensor-core phases from 32 streams on)
    assert outs[True][1] == 1 and outs[False][1] > 10, (outs[True][1], outs[False][1])
