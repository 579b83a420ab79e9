"""Per-kernel parity on the B200: every C-ABI entry point against the same arithmetic in torch fp32 / the oracle.
Tolerances: these kernels are exact fp32 (FMA accumulation), so only summation-order noise is allowed."""
import pytest
import torch
import torch.nn.functional as F

from oracle import separator_ref as R

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from puresound_b200 import ops as o

    o.require_device()
    return o


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (scale * (2 * torch.rand(*shape, generator=g) - 1)).to(DEV)


def close(a, b, tol=1e-5):
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    ref = max(1.0, b.abs().max().item())
    assert err <= tol * ref, f"max abs err {err} (ref scale {ref})"


def act_t(x, act, slope=None):
    return {0: lambda v: v, 1: lambda v: F.prelu(v, slope), 2: torch.relu, 3: torch.tanh, 4: torch.sigmoid}[act](x)


@pytest.mark.parametrize("B,Rr,M,K", [(2, 37, 20, 13), (1, 300, 130, 64), (3, 129, 257, 100), (2, 128, 128, 16), (1, 5, 512, 512)])
def test_gemm_plain_bias_residual_act(ops, B, Rr, M, K):
    x, w, b, res = rnd(B, Rr, K, seed=1), rnd(M, K, seed=2), rnd(M, seed=3), rnd(B, Rr, M, seed=4)
    bb = rnd(B, M, seed=5)
    y, _ = ops.linear(x, w, bias=b, bias_batch=bb, epi_act=ops.ACT_RELU, residual=res)
    ref = torch.relu(x @ w.t() + b + bb.unsqueeze(1)) + res
    close(y, ref)


@pytest.mark.parametrize("B,Rr,M,K", [(1, 1, 512, 512), (1, 3, 320, 512), (2, 4, 100, 64), (1, 8, 512, 320), (8, 1, 192, 1024), (1, 19, 512, 512),
                                       (4, 8, 256, 512), (1, 32, 512, 1024)])
@pytest.mark.parametrize("mode", ["none", "affine", "mask"])
def test_gemm_few_rows_kernel(ops, B, Rr, M, K, mode):
    """batch * rows <= 32, eight per pass (a few streams' per-hop 1x1 convs, heads on pooled embeddings): warp-per-output-channel kernel
    with the operand rows in shared memory; every prologue it takes, bias / per-item bias / activation / residual."""
    x, w, b, res = rnd(B, Rr, K, seed=1), rnd(M, K, seed=2, scale=0.2), rnd(M, seed=3), rnd(B, Rr, M, seed=4)
    bb, slope = rnd(B, M, seed=5), torch.tensor([0.2], device=DEV)
    if mode == "affine":
        sc, sh = rnd(B, K, seed=6) + 1.5, rnd(B, K, seed=7)
        pro = ops.Prologue(ops.PRO_AFFINE, ops.ACT_PRELU, sc, sh, K, None, slope)
        xin = F.prelu(x * sc.unsqueeze(1) + sh.unsqueeze(1), slope)
    elif mode == "mask":
        mk = rnd(B, Rr, K, seed=8, scale=2)
        pro = ops.Prologue(ops.PRO_MASK, ops.ACT_SIGMOID, x2=mk)
        xin = x * torch.sigmoid(mk)
    else:
        pro, xin = ops.NO_PRO, x
    y, _ = ops.linear(x, w, pro=pro, bias=b, bias_batch=bb, epi_act=ops.ACT_PRELU, epi_slope=slope, residual=res, backend=ops.GEMM_SIMT)
    ref = F.prelu((xin.double() @ w.double().t() + b.double() + bb.double().unsqueeze(1)).float(), slope) + res
    close(y, ref, 1e-5)


def test_gemm_framed_view_is_conv1d(ops):
    """row stride < K: the waveform read in place as overlapping frames == F.conv1d (encoder.py:50-56)."""
    wav, w = rnd(3, 1000, seed=1), rnd(24, 32, seed=2)
    T = (1000 - 32) // 16 + 1
    y, _ = ops.gemm(wav, w, batch=3, rows=T, M=24, K=32, x_batch_stride=1000, x_row_stride=16, w_row_stride=32)
    close(y, F.conv1d(wav.unsqueeze(1), w.unsqueeze(1), stride=16).transpose(1, 2))
    # odd hop: exercises the unaligned scalar loader
    T = (1000 - 30) // 7 + 1
    w = rnd(5, 30, seed=3)
    y, _ = ops.gemm(wav, w, batch=3, rows=T, M=5, K=30, x_batch_stride=1000, x_row_stride=7, w_row_stride=30)
    close(y, F.conv1d(wav.unsqueeze(1), w.unsqueeze(1), stride=7).transpose(1, 2))


@pytest.mark.parametrize("act", [0, 1, 3])
def test_gemm_affine_prologue_and_stats(ops, act):
    B, Rr, M, K = 2, 200, 48, 40
    x, w = rnd(B, Rr, K, seed=1, scale=3), rnd(M, K, seed=2)
    sc, sh, slope = rnd(B, K, seed=3) + 1.5, rnd(B, K, seed=4), torch.tensor([0.2], device=DEV)
    y, part = ops.linear(x, w, pro=ops.Prologue(ops.PRO_AFFINE, act, sc, sh, K, None, slope), want_stats=True)
    xin = act_t(x * sc.unsqueeze(1) + sh.unsqueeze(1), act, slope)
    ref = xin @ w.t()
    close(y, ref)
    gamma, beta = rnd(M, seed=5) + 1.5, rnd(M, seed=6)
    scale, shift = ops.stats_finalize(part, gamma, beta, 1e-8, M)
    mu = ref.mean(dim=(1, 2), keepdim=True)
    var = (ref - mu).pow(2).mean(dim=(1, 2), keepdim=True)
    rstd = 1 / torch.sqrt(var + 1e-8)
    close(scale, gamma * rstd.view(B, 1))
    close(shift, beta - (mu * rstd).view(B, 1) * gamma, 2e-5)
    # batch-independent affine (bN1d fold): stride 0
    y0, _ = ops.linear(x, w, pro=ops.Prologue(ops.PRO_AFFINE, act, sc[0].contiguous(), sh[0].contiguous(), 0, None, slope))
    close(y0, act_t(x * sc[0] + sh[0], act, slope) @ w.t())


def test_gemm_rownorm_and_mask_prologues(ops):
    B, Rr, M, K = 2, 77, 33, 24
    x, w = rnd(B, Rr, K, seed=1, scale=2) + 0.3, rnd(M, K, seed=2)
    g, b, slope = rnd(K, seed=3) + 1.5, rnd(K, seed=4), torch.tensor([0.3], device=DEV)
    rs = ops.rowstats(x, 1e-8)
    close(rs[..., 0], x.mean(-1))
    close(rs[..., 1], 1 / torch.sqrt(x.var(-1, unbiased=False) + 1e-8))
    y, _ = ops.linear(x, w, pro=ops.Prologue(ops.PRO_ROWNORM, ops.ACT_PRELU, g, b, 0, rs, slope))
    xin = F.prelu(F.layer_norm(x, (K,), g, b, 1e-8), slope)
    close(y, xin @ w.t())
    m = rnd(B, Rr, K, seed=5)
    y, _ = ops.linear(x, w, pro=ops.Prologue(ops.PRO_MASK, ops.ACT_RELU, x2=m))
    close(y, (x * torch.relu(m)) @ w.t())


def test_gemm_embedding_columns_view(ops):
    """W given as a row-strided column slice (the W_in[:, C:] embedding fold of conv_tasnet.py:80-83)."""
    H, Cc, E, N = 24, 16, 6, 3
    w_full, e = rnd(H, Cc + E, seed=1), rnd(N, E, seed=2)
    y, _ = ops.gemm(e, w_full[:, Cc:], batch=1, rows=N, M=H, K=E, x_batch_stride=0, x_row_stride=E, w_row_stride=Cc + E)
    close(y[0], e @ w_full[:, Cc:].t())


@pytest.mark.parametrize("C,T,P,d,causal", [(512, 300, 3, 8, False), (24, 50, 3, 2, True), (18, 41, 3, 1, False), (64, 100, 5, 3, True), (256, 70, 3, 64, False)])
@pytest.mark.parametrize("mode", ["none", "affine", "rownorm"])
def test_dwconv(ops, C, T, P, d, causal, mode):
    B = 2
    x, w, b = rnd(B, T, C, seed=1, scale=2), rnd(C, P, seed=2), rnd(C, seed=3)
    slope = torch.tensor([0.25], device=DEV)
    if mode == "none":
        pro, xin = ops.NO_PRO, x
    elif mode == "affine":
        sc, sh = rnd(B, C, seed=4) + 1.5, rnd(B, C, seed=5)
        pro = ops.Prologue(ops.PRO_AFFINE, ops.ACT_PRELU, sc, sh, C, None, slope)
        xin = F.prelu(x * sc.unsqueeze(1) + sh.unsqueeze(1), slope)
    else:
        g, bt = rnd(C, seed=6) + 1.5, rnd(C, seed=7)
        pro = ops.Prologue(ops.PRO_ROWNORM, ops.ACT_PRELU, g, bt, 0, ops.rowstats(x, 1e-8), slope)
        xin = F.prelu(F.layer_norm(x, (C,), g, bt, 1e-8), slope)
    y, part = ops.dwconv(x, w, b, P, d, causal, pro, want_stats=True)
    pad = (P - 1) * d if causal else ((P - 1) // 2) * d
    ref = F.conv1d(xin.transpose(1, 2), w.unsqueeze(1), b, dilation=d, padding=pad, groups=C)
    if causal:
        ref = ref[..., :-pad]
    ref = ref.transpose(1, 2)
    close(y, ref)
    scale, shift = ops.stats_finalize(part, None, None, 1e-8, C)
    mu = ref.mean(dim=(1, 2))
    rstd = 1 / torch.sqrt(ref.var(dim=(1, 2), unbiased=False) + 1e-8)
    close(scale[:, 0], rstd)
    close(shift[:, 0], -mu * rstd, 2e-5)
    # fused finalize (tiled kernel: last CTA of an item merges; streaming kernel: follow-up launch inside the library):
    # the same folded affine as the stand-alone merge, twice in a row (the per-item counters must come back to zero)
    gamma, beta = rnd(C, seed=8) + 1.5, rnd(C, seed=9)
    sc_ref, sh_ref = ops.stats_finalize(part, gamma, beta, 1e-8, C)
    for _ in range(2):
        y2, fa = ops.dwconv(x, w, b, P, d, causal, pro, want_stats=True, fin=(gamma, beta, 1e-8))
        # (shapes the TMA sliding-window kernel serves take the tiled kernel when the finalize is fused: same taps in the
        # same order, but another partition of the statistics partials - equal to rounding, not bit for bit)
        assert torch.equal(y2, y)
        close(fa.scale, sc_ref, 1e-6)
        close(fa.shift, sh_ref, 1e-5)


def test_rownorm_layernorm_residual(ops):
    x, res, w, b = rnd(5, 9, 128, seed=1, scale=3), rnd(5, 9, 128, seed=2), rnd(128, seed=3) + 1.5, rnd(128, seed=4)
    close(ops.rownorm(x, w, b, 1e-5, res=res), res + F.layer_norm(x, (128,), w, b, 1e-5))
    slope = torch.tensor([0.1], device=DEV)
    close(ops.rownorm(x, w, b, 1e-8, act=ops.ACT_PRELU, slope=slope), F.prelu(F.layer_norm(x, (128,), w, b, 1e-8), slope))


def test_bn_fold(ops):
    w, b, rm, rv = rnd(40, seed=1) + 1.5, rnd(40, seed=2), rnd(40, seed=3), rnd(40, seed=4) + 1.5
    sc, sh = ops.bn_fold(w, b, rm, rv, 1e-5)
    x = rnd(2, 40, 30, seed=5)
    close(x * sc.view(1, -1, 1) + sh.view(1, -1, 1), F.batch_norm(x, rm, rv, w, b, False, 0.0, 1e-5))


@pytest.mark.parametrize("win,hop,T", [(32, 16, 50), (64, 16, 10), (320, 160, 7), (30, 7, 20)])
def test_ola(ops, win, hop, T):
    fr = rnd(3, T, win, seed=1)
    out_len = (T - 1) * hop + win
    ref = F.fold(fr.transpose(1, 2), (1, out_len), kernel_size=(1, win), stride=hop).flatten(1)
    close(ops.ola(fr, hop, None, 0), ref)
    close(ops.ola(fr * 3, hop, None, 1), torch.clamp(ref * 3, -1, 1))
    close(ops.ola(fr, hop, None, 2), torch.sigmoid(ref))
    wsum = F.fold(torch.ones_like(fr[:1]).transpose(1, 2) * 0.7, (1, out_len), kernel_size=(1, win), stride=hop).flatten()
    wsum[0] = 0.0
    exp = ref.clone()
    exp[:, 1:] = exp[:, 1:] / wsum[1:]
    close(ops.ola(fr, hop, wsum, 0), exp)


def test_mask_magnitude_l2(ops):
    f, m = rnd(4, 11, 32, seed=1), rnd(4, 11, 32, seed=2)
    close(ops.mask_apply(f, m, ops.ACT_RELU, False), f * torch.relu(m))
    close(ops.mask_apply(f, m, ops.ACT_SIGMOID, False), f * torch.sigmoid(m))
    ref = R.apply_tf_masks(f.transpose(1, 2), m.transpose(1, 2), "complex", "complex").transpose(1, 2)
    close(ops.mask_apply(f, m, ops.ACT_NONE, True), ref)
    close(ops.magnitude(f, False, False), R.magnitude(f.transpose(1, 2), False).transpose(1, 2))
    close(ops.magnitude(f, True, True), R.magnitude(f.transpose(1, 2), True, True).transpose(1, 2))
    e = rnd(5, 192, seed=3)
    close(ops.l2normalize(e), F.normalize(e, p=2, dim=1))


def test_asp_pool(ops):
    x, lg = rnd(3, 200, 70, seed=1), rnd(3, 200, 70, seed=2, scale=4)
    w = torch.softmax(lg, dim=1)
    mean = (w * x).sum(1)
    std = torch.sqrt((w * (x - mean.unsqueeze(1)).pow(2)).sum(1).clamp(1e-12))
    close(ops.asp_pool(x, lg), torch.cat([mean, std], 1))


@pytest.mark.parametrize("T,K", [(103, 10), (57, 8), (100, 20), (999, 100)])
def test_segment_merge_bit_exact(ops, T, K):
    from puresound_b200.nnet.lobe.trivial import overlap_geometry

    x = rnd(2, T, 12, seed=1)
    seg_ref, rest = R.split_overlap(x.cpu().transpose(1, 2), K)
    r, S = overlap_geometry(T, K)
    seg = ops.segment(x, K, S, True)
    assert torch.equal(seg.cpu(), seg_ref)
    y = rnd(2, S, K, 12, seed=2)
    assert torch.equal(ops.merge(y, T, True).cpu(), R.merge_overlap(y.cpu(), rest).transpose(1, 2))
    assert torch.equal(ops.merge(seg, T, True), x)  # reference pin test/test_lobe.py:49-54
    # no-overlap: zero-padded reshape incl. the whole extra segment when T % K == 0
    S2 = (T + (K - T % K)) // K
    seg2 = ops.segment(x, K, S2, False)
    ref2, _ = R._dprnn_segment(x.cpu().transpose(1, 2), K, False)
    assert torch.equal(seg2.cpu(), ref2)
    assert torch.equal(ops.merge(seg2, T, False), x)


@pytest.mark.parametrize("M,K,hop,relu", [(512, 32, 16, False), (128, 32, 16, True), (96, 20, 10, False), (640, 64, 32, True)])
def test_short_k_filterbank_kernel(ops, M, K, hop, relu):
    """FreeEncDec analysis (lobe/encoder.py:50-56,71-83) at short windows goes through the register-resident filterbank
    kernel: framed rows read in place, optional ReLU (output_active) and bias, ragged last CTA, channel tails."""
    N, L = 3, 16000 + 7 * hop
    wav, w, bias = rnd(N, L, seed=1), rnd(M, K, seed=2, scale=0.2), rnd(M, seed=3)
    T = (L - K) // hop + 1
    y, _ = ops.gemm(wav, w, batch=N, rows=T, M=M, K=K, x_batch_stride=L, x_row_stride=hop, w_row_stride=K, bias=bias,
                    epi_act=ops.ACT_RELU if relu else ops.ACT_NONE, backend=ops.GEMM_SIMT)
    ref = torch.nn.functional.conv1d(wav.unsqueeze(1).double(), w.unsqueeze(1).double(), bias.double(), stride=hop).transpose(1, 2)
    if relu:
        ref = torch.relu(ref)
    close(y.double(), ref, 1e-6)


@pytest.mark.parametrize("M,K,hop,relu", [(2, 384, 128, False), (2, 256, 128, True), (5, 100, 100, False), (8, 1024, 64, False)])
def test_thin_output_gemm_kernel(ops, M, K, hop, relu):
    """M <= 8 output channels (the U-Net shell's output layer: ConvTranspose2d to 2 channels as framed GEMM rows,
    unet.py:154-170) go through the warp-per-row kernel: overlapping rows read in place, bias, activation, strided output."""
    N, L = 3, K + hop * 2999
    x, w, bias = rnd(N, L, seed=1), rnd(M, K, seed=2, scale=0.2), rnd(M, seed=3)
    T = (L - K) // hop + 1
    out = torch.full((N, T, 2 * M + 3), 7.0, device=DEV)
    ops.gemm(x, w, batch=N, rows=T, M=M, K=K, x_batch_stride=L, x_row_stride=hop, w_row_stride=K, bias=bias,
             epi_act=ops.ACT_RELU if relu else ops.ACT_NONE, backend=ops.GEMM_SIMT, out=out.view(-1)[M:], y_strides=(T * (2 * M + 3), 2 * M + 3))
    ref = torch.nn.functional.conv1d(x.unsqueeze(1).double(), w.unsqueeze(1).double(), bias.double(), stride=hop).transpose(1, 2)
    if relu:
        ref = torch.relu(ref)
    close(out[:, :, M:2 * M].double(), ref, 2e-6)
    assert (out[:, :, :M] == 7.0).all() and (out[:, :, 2 * M:] == 7.0).all()  # only its own columns are written


@pytest.mark.parametrize("H,D,N,S,K", [
    (128, 1, 2, 37, 50), (128, 2, 2, 37, 50),   # 74 sequences: one wave whatever the CTA size -> 32 per CTA
    (64, 1, 2, 37, 50), (64, 2, 1, 21, 30),     # veve_dprnn_v0_causal's hidden size: units 64..127 are zero padding
    (96, 2, 1, 19, 12), (32, 1, 3, 11, 9),
    (128, 2, 4, 601, 6),                         # 2404 sequences x 2 directions: 48 per CTA (second chunk half empty)
    (128, 1, 4, 2250, 3),                        # 9000 sequences: one wave only with 64 per CTA
    (64, 1, 4, 751, 5),                          # 3004 sequences: 32 per CTA
])
def test_lstm_tensor_core_path(ops, H, D, N, S, K):
    """H <= 128: W_hh resident as bf16 hi (tensor memory) + lo (shared memory), gates on tcgen05 with the 3xBF16 split.
    Same oracle (step-by-step LSTM cell, dprnn.py:67-103 via nn.LSTM) as the exact-fp32 kernel; ragged last CTA, every
    sequences-per-CTA choice of the launcher, both addressing modes, initial and final state."""
    C = 16
    sd = {}
    for s in ["", "_reverse"][:D]:
        sd[f"weight_ih_l0{s}"], sd[f"weight_hh_l0{s}"] = rnd(4 * H, C, seed=1, scale=0.3).cpu(), rnd(4 * H, H, seed=2, scale=0.15).cpu()
        sd[f"bias_ih_l0{s}"], sd[f"bias_hh_l0{s}"] = rnd(4 * H, seed=3, scale=0.3).cpu(), rnd(4 * H, seed=4, scale=0.3).cpu()
    sfx = ["", "_reverse"][:D]
    w_ih = torch.cat([sd[f"weight_ih_l0{s}"] for s in sfx]).to(DEV)
    b = torch.cat([sd[f"bias_ih_l0{s}"] + sd[f"bias_hh_l0{s}"] for s in sfx]).to(DEV)
    w_hh_t = torch.stack([sd[f"weight_hh_l0{s}"].t().contiguous() for s in sfx]).to(DEV)
    pk = ops.lstm_pack_weights(w_hh_t, H, D)
    assert pk is not None and pk.numel() == D * 4 * 128 * 128 * 4  # the on-chip layout is always 128 units wide
    x = rnd(N, S, K, C, seed=5)
    P = N * S * K
    gx, _ = ops.linear(x.view(1, P, C), w_ih, bias=b)
    gx = gx.view(P, D * 4 * H)
    out, st = ops.lstm(gx, w_hh_t, n_seq=N * S, L=K, H=H, D=D, inner=1, outer_stride=K, inner_stride=0, step_stride=1,
                       want_state=True, w_packed=pk)
    ref, (hn, cn) = R.lstm(sd, "", x.cpu().view(N * S, K, C), D == 2, None, fast=True)
    close(out.view(N * S, K, D * H).cpu(), ref, 5e-5)
    close(st[0].cpu(), hn, 5e-5)
    close(st[1].cpu(), cn, 2e-4)
    # and it agrees with the exact-fp32 CUDA-core kernel
    out32, _ = ops.lstm(gx, w_hh_t, n_seq=N * S, L=K, H=H, D=D, inner=1, outer_stride=K, inner_stride=0, step_stride=1)
    close(out, out32, 5e-5)
    # gx with the rows of W_ih permuted to [dir][unit][gate] (one 16-byte load per sequence) gives the same result
    gxi = gx.view(P, D, 4, H).permute(0, 1, 3, 2).reshape(P, D * 4 * H).contiguous()
    outi, _ = ops.lstm(gxi, w_hh_t, n_seq=N * S, L=K, H=H, D=D, inner=1, outer_stride=K, inner_stride=0, step_stride=1,
                       w_packed=pk, gx_interleaved=True)
    assert torch.equal(outi, out)
    h0, c0 = rnd(D, N * K, H, seed=6), rnd(D, N * K, H, seed=7)
    out, st = ops.lstm(gx, w_hh_t, n_seq=N * K, L=S, H=H, D=D, inner=K, outer_stride=S * K, inner_stride=1, step_stride=K,
                       h0=h0, c0=c0, want_state=True, w_packed=pk)
    xi = x.cpu().permute(0, 2, 1, 3).reshape(N * K, S, C)
    ref, (hn, cn) = R.lstm(sd, "", xi, D == 2, (h0.cpu(), c0.cpu()), fast=True)
    got = out.view(N, S, K, D * H).permute(0, 2, 1, 3).reshape(N * K, S, D * H).cpu()
    close(got, ref, 5e-5)
    close(st[0].cpu(), hn, 5e-5)
    close(st[1].cpu(), cn, 2e-4)


@pytest.mark.parametrize("H,D", [(12, 1), (12, 2), (64, 1), (128, 2), (256, 1)])
def test_lstm_intra_and_inter_addressing(ops, H, D):
    N, S, K, C = 2, 5, 7, 16
    sd = {}
    for s in ["", "_reverse"][:D]:
        sd[f"weight_ih_l0{s}"], sd[f"weight_hh_l0{s}"] = rnd(4 * H, C, seed=1, scale=0.3).cpu(), rnd(4 * H, H, seed=2, scale=0.3).cpu()
        sd[f"bias_ih_l0{s}"], sd[f"bias_hh_l0{s}"] = rnd(4 * H, seed=3, scale=0.3).cpu(), rnd(4 * H, seed=4, scale=0.3).cpu()
    sfx = ["", "_reverse"][:D]
    w_ih = torch.cat([sd[f"weight_ih_l0{s}"] for s in sfx]).to(DEV)
    b = torch.cat([sd[f"bias_ih_l0{s}"] + sd[f"bias_hh_l0{s}"] for s in sfx]).to(DEV)
    w_hh_t = torch.stack([sd[f"weight_hh_l0{s}"].t().contiguous() for s in sfx]).to(DEV)
    x = rnd(N, S, K, C, seed=5)
    P = N * S * K
    gx, _ = ops.linear(x.view(1, P, C), w_ih, bias=b)
    gx = gx.view(P, D * 4 * H)
    # intra: N*S sequences over K
    out, st = ops.lstm(gx, w_hh_t, n_seq=N * S, L=K, H=H, D=D, inner=1, outer_stride=K, inner_stride=0, step_stride=1, want_state=True)
    ref, (hn, cn) = R.lstm(sd, "", x.cpu().view(N * S, K, C), D == 2, None, fast=False)
    close(out.view(N * S, K, D * H).cpu(), ref, 2e-5)
    close(st[0].cpu(), hn, 2e-5)
    close(st[1].cpu(), cn, 2e-5)
    # inter: N*K sequences over S with initial state, no permute of the data
    h0, c0 = rnd(D, N * K, H, seed=6), rnd(D, N * K, H, seed=7)
    out, st = ops.lstm(gx, w_hh_t, n_seq=N * K, L=S, H=H, D=D, inner=K, outer_stride=S * K, inner_stride=1, step_stride=K,
                       h0=h0, c0=c0, want_state=True)
    xi = x.cpu().permute(0, 2, 1, 3).reshape(N * K, S, C)
    ref, (hn, cn) = R.lstm(sd, "", xi, D == 2, (h0.cpu(), c0.cpu()), fast=True)
    got = out.view(N, S, K, D * H).permute(0, 2, 1, 3).reshape(N * K, S, D * H).cpu()
    close(got, ref, 2e-5)
    close(st[0].cpu(), hn, 2e-5)
    close(st[1].cpu(), cn, 2e-5)


@pytest.mark.parametrize("H,D,N,L", [(12, 2, 37, 9), (40, 1, 300, 6), (256, 1, 70, 5), (256, 2, 1300, 4), (200, 1, 9, 3)])
def test_lstm_cuda_core_packed_weights(ops, H, D, N, L):
    """Sizes the tensor-core recurrence does not serve (SkiM's H = 256, small test sizes): ps_lstm_pack_weights builds the
    gate-minor fp32 image [D][k][unit][4] (one 16-byte weight load per k, prefetched two chunks ahead); 8 or 16 sequences per
    thread (N = 1300 at H = 256 takes the 16-wide variant); with and without interleaved gx rows, initial / final states.
    Bit-identical to the unpacked kernel (same fp32 operations in the same order)."""
    C = 16
    sd = {}
    for s in ["", "_reverse"][:D]:
        sd[f"weight_ih_l0{s}"], sd[f"weight_hh_l0{s}"] = rnd(4 * H, C, seed=1, scale=0.3).cpu(), rnd(4 * H, H, seed=2, scale=0.1).cpu()
        sd[f"bias_ih_l0{s}"], sd[f"bias_hh_l0{s}"] = rnd(4 * H, seed=3, scale=0.3).cpu(), rnd(4 * H, seed=4, scale=0.3).cpu()
    sfx = ["", "_reverse"][:D]
    w_ih = torch.cat([sd[f"weight_ih_l0{s}"] for s in sfx]).to(DEV)
    b = torch.cat([sd[f"bias_ih_l0{s}"] + sd[f"bias_hh_l0{s}"] for s in sfx]).to(DEV)
    w_hh_t = torch.stack([sd[f"weight_hh_l0{s}"].t().contiguous() for s in sfx]).to(DEV)
    pk = ops.lstm_pack_weights(w_hh_t, H, D)
    assert pk is not None and pk.numel() == D * H * 4 * H * 4
    x = rnd(N, L, C, seed=5)
    P = N * L
    gx, _ = ops.linear(x.view(1, P, C), w_ih, bias=b)
    gx = gx.view(P, D * 4 * H)
    h0, c0 = rnd(D, N, H, seed=6), rnd(D, N, H, seed=7)
    geo = dict(n_seq=N, L=L, H=H, D=D, inner=1, outer_stride=L, inner_stride=0, step_stride=1)
    ref, (hn, cn) = R.lstm(sd, "", x.cpu(), D == 2, (h0.cpu(), c0.cpu()), fast=True)
    plain, st0 = ops.lstm(gx, w_hh_t, h0=h0, c0=c0, want_state=True, **geo)
    out, st = ops.lstm(gx, w_hh_t, h0=h0, c0=c0, want_state=True, w_packed=pk, **geo)
    close(out.view(N, L, D * H).cpu(), ref, 2e-5)
    close(st[0].cpu(), hn, 2e-5)
    close(st[1].cpu(), cn, 2e-5)
    if not (H == 256 and N > 1184):  # (the 16-wide variant sums in the same order too, but keep the claim to the like-for-like pair)
        assert torch.equal(out, plain) and torch.equal(st[0], st0[0]) and torch.equal(st[1], st0[1])
    gxi = gx.view(P, D, 4, H).permute(0, 1, 3, 2).reshape(P, D * 4 * H).contiguous()
    outi, sti = ops.lstm(gxi, w_hh_t, h0=h0, c0=c0, want_state=True, w_packed=pk, gx_interleaved=True, **geo)
    assert torch.equal(outi, out) and torch.equal(sti[1], st[1])


def test_film_combine_and_transpose(ops):
    sb, xn = rnd(50, 32, seed=1), rnd(50, 16, seed=2)
    close(ops.film_combine(sb, xn), sb[:, :16] * xn + sb[:, 16:])
    x = rnd(3, 45, 70, seed=3)
    assert torch.equal(ops.transpose(x), x.transpose(1, 2).contiguous())


def test_nan_inf_propagate(ops):
    """_verbose() (base_nn.py:740-777) finds look-ahead by where NaNs appear: kernels must not flush them."""
    x = rnd(1, 40, 16, seed=1)
    x[0, 20:, :] = float("inf")
    w = rnd(8, 16, seed=2)
    y, part = ops.linear(x, w, want_stats=True)
    assert torch.isfinite(y[0, :20]).all() and not torch.isfinite(y[0, 20:]).any()
    sc, sh = ops.stats_finalize(part, None, None, 1e-8, 8)
    assert not torch.isfinite(sh).any()
    fr = torch.full((1, 4, 32), float("nan"), device=DEV)
    assert torch.isnan(ops.ola(fr, 16, None, 1)).all()
    assert torch.isnan(ops.mask_apply(fr, fr, ops.ACT_RELU, False)).all()


def test_host_tensors_are_rejected(ops):
    with pytest.raises(TypeError):
        ops.linear(torch.zeros(1, 4, 4), torch.zeros(4, 4))


@pytest.mark.parametrize("B,C,T,d,causal,mode", [
    (2, 64, 3999, 1, False, "affine"), (2, 64, 3999, 128, False, "affine"), (3, 512, 1000, 16, False, "affine"), (2, 32, 700, 64, True, "affine"),
    (1, 512, 3999, 32, False, "affine"), (2, 96, 64, 2, False, "none"), (5, 128, 777, 160, True, "none"), (64, 32, 500, 8, False, "affine")])
def test_dwconv_tma_sliding_window(ops, B, C, T, d, causal, mode):
    """dwconv_tma_kernel (P = 3, C % 32 == 0): rows enter a shared-memory ring by tensor-map TMA, are normalised + PReLU'd in
    place once and serve their three taps from the ring.  Runs shorter than / equal to / longer than the item, the largest
    dilations of the TCN stack (halo up to 320 frames, 7-8 ring slots), causal and centred padding (zero AFTER the prologue),
    several runs per item (halo re-read where runs meet), against F.conv1d; the statistics of the whole item from the per-run
    partials; and bit-equality with dwconv_tile_kernel's taps order is NOT required (only the 1e-5 tolerance)."""
    P = 3
    x, w, b = rnd(B, T, C, seed=1, scale=2), rnd(C, P, seed=2), rnd(C, seed=3)
    slope = torch.tensor([0.25], device=DEV)
    if mode == "none":
        pro, xin = ops.NO_PRO, x
    else:
        sc, sh = rnd(B, C, seed=4) + 1.5, rnd(B, C, seed=5)
        pro = ops.Prologue(ops.PRO_AFFINE, ops.ACT_PRELU, sc, sh, C, None, slope)
        xin = F.prelu(x * sc.unsqueeze(1) + sh.unsqueeze(1), slope)
    y, part = ops.dwconv(x, w, b, P, d, causal, pro, want_stats=True)
    pad = (P - 1) * d if causal else ((P - 1) // 2) * d
    ref = F.conv1d(xin.transpose(1, 2), w.unsqueeze(1), b, dilation=d, padding=pad, groups=C)
    if causal:
        ref = ref[..., :-pad]
    ref = ref.transpose(1, 2)
    close(y, ref)
    scale, shift = ops.stats_finalize(part, None, None, 1e-8, C)
    mu = ref.mean(dim=(1, 2))
    rstd = 1 / torch.sqrt(ref.var(dim=(1, 2), unbiased=False) + 1e-8)
    close(scale[:, 0], rstd)
    close(shift[:, 0], -mu * rstd, 2e-5)
    # NaN / Inf stay where the taps reach (the reference's look-ahead probe, base_nn.py:740-777)
    x2 = x.clone()
    x2[0, T // 2:, :] = float("inf")
    y2, _ = ops.dwconv(x2, w, b, P, d, causal, pro)
    reach = 0 if causal else d
    first_bad = max(T // 2 - reach, 0)
    assert torch.isfinite(y2[0, :first_bad]).all() and not torch.isfinite(y2[0, T // 2:]).any()
    assert torch.equal(y2[1:], y[1:])
