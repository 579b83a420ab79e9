"""tcgen05/TMEM GEMM (3xBF16 split, fp32 TMEM accumulate) against an fp64 reference and against the exact-fp32
CUDA-core kernel.  Tolerance: ~2^-17 relative per product (hi*hi + hi*lo + lo*hi), i.e. 5e-5 of the output scale —
30x tighter than TF32 would give; end-to-end waveform parity (1e-3) is checked in test_gpu_full.py."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 5e-5


@pytest.fixture(scope="module")
def ops():
    from puresound_b200 import ops as o

    o.require_device()
    return o


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (scale * (2 * torch.rand(*shape, generator=g) - 1)).to(DEV)


def check(y, ref64, tol=TOL):
    err = (y.double() - ref64).abs().max().item()
    scale = max(1.0, ref64.abs().max().item())
    assert err <= tol * scale, f"max abs err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("B,Rr,M,K", [(1, 128, 256, 64), (2, 300, 512, 512), (3, 129, 256, 128), (1, 1000, 512, 256), (5, 77, 512, 64)])
def test_tc_plain(ops, B, Rr, M, K):
    x, w = rnd(B, Rr, K, seed=1, scale=2), rnd(M, K, seed=2, scale=0.1)
    pk = ops.pack_weights(w, M, K, K)
    assert pk is not None and pk.numel() == M * K * 4
    y, _ = ops.linear(x, w, w_packed=pk, backend=ops.GEMM_TCGEN05)
    check(y, x.double() @ w.double().t())
    y2, _ = ops.linear(x, w, backend=ops.GEMM_SIMT)
    check(y2, x.double() @ w.double().t(), 1e-5)


@pytest.mark.parametrize("B,Rr,M,K", [(2, 300, 128, 256), (1, 200, 384, 64), (1, 150, 640, 128)])
def test_tc_channel_counts_padded_to_256(ops, B, Rr, M, K):
    """M = 128 (the DPRNN projections), 384, 640: the CTA-pair kernel works on whole 256-channel blocks; the packed image
    pads the missing rows with zeros and the epilogue skips them (output, bias, statistics)."""
    x, w, bias = rnd(B, Rr, K, seed=1, scale=2), rnd(M, K, seed=2, scale=0.1), rnd(M, seed=3)
    pk = ops.pack_weights(w, M, K, K)
    assert pk is not None
    y, part = ops.linear(x, w, bias=bias, want_stats=True, w_packed=pk, backend=ops.GEMM_TCGEN05)
    ref = x.double() @ w.double().t() + bias.double()
    check(y, ref)
    scale, shift = ops.stats_finalize(part, None, None, 1e-8, M)
    rstd = 1 / torch.sqrt(ref.var(dim=(1, 2), unbiased=False) + 1e-8)
    assert (scale[:, 0].double() - rstd).abs().max() <= 1e-5 * rstd.abs().max()


@pytest.mark.parametrize("act", ["none", "relu", "sigmoid"])
def test_tc_mask_prologue_small_m(ops, act):
    """The decoder GEMM of FreeEncDec: x * act(mask) applied on load (base_nn.py:81-95,146-159), M = win = 32 output
    samples per frame (zero-padded to one 256-channel block, lane quarters 1-3 skipped)."""
    B, Rr, M, K = 2, 333, 32, 512
    x, mk, w = rnd(B, Rr, K, seed=1, scale=2), rnd(B, Rr, K, seed=2, scale=2), rnd(M, K, seed=3, scale=0.1)
    code = {"none": ops.ACT_NONE, "relu": ops.ACT_RELU, "sigmoid": ops.ACT_SIGMOID}[act]
    f = {"none": lambda t: t, "relu": torch.relu, "sigmoid": torch.sigmoid}[act]
    pk = ops.pack_weights(w, M, K, K)
    assert pk is not None
    y, _ = ops.linear(x, w, pro=ops.Prologue(ops.PRO_MASK, code, x2=mk), w_packed=pk, backend=ops.GEMM_TCGEN05)
    check(y, (x.double() * f(mk.double())) @ w.double().t())


def test_tc_overlapping_rows_framed_view(ops):
    """Framed analysis GEMM: rows are overlapping windows of a waveform (row stride = hop < K), read in place."""
    N, L, win, hop, M = 2, 16000, 512, 128, 512
    wav, w = rnd(N, L, seed=1), rnd(M, win, seed=2, scale=0.05)
    T = (L - win) // hop + 1
    pk = ops.pack_weights(w, M, win, win)
    y, _ = ops.gemm(wav, w, batch=N, rows=T, M=M, K=win, x_batch_stride=L, x_row_stride=hop, w_row_stride=win,
                    w_packed=pk, backend=ops.GEMM_TCGEN05)
    frames = wav.unfold(1, win, hop)  # [N, T, win]
    check(y, frames.double() @ w.double().t())


@pytest.mark.parametrize("N,L,M,relu", [(1, 16 * 127 + 32, 512, False), (3, 16000, 512, True), (2, 64000, 128, True), (5, 4000, 256, False),
                                        (64, 64000, 512, False)])
def test_tc_short_window_k32(ops, N, L, M, relu):
    """The learned encoder of FreeEncDec (lobe/encoder.py:50-56,71-83) at win = 32, hop = 16: K = 32 is ONE shared-memory
    stage per tile, the two producer halves take alternate tiles.  One tile, odd / even tile counts per CTA pair, M = 128
    (DPRNN front-end) and the full cfg2 shape (more tiles than CTA pairs)."""
    win, hop = 32, 16
    wav, w = rnd(N, L, seed=1), rnd(M, win, seed=2, scale=0.2)
    T = (L - win) // hop + 1
    pk = ops.pack_weights(w, M, win, win)
    assert pk is not None
    y, _ = ops.gemm(wav, w, batch=N, rows=T, M=M, K=win, x_batch_stride=L, x_row_stride=hop, w_row_stride=win,
                    w_packed=pk, backend=ops.GEMM_TCGEN05, epi_act=ops.ACT_RELU if relu else ops.ACT_NONE)
    y2, _ = ops.gemm(wav, w, batch=N, rows=T, M=M, K=win, x_batch_stride=L, x_row_stride=hop, w_row_stride=win,
                     backend=ops.GEMM_SIMT, epi_act=ops.ACT_RELU if relu else ops.ACT_NONE)
    for n0 in range(0, N, 8):  # fp64 reference in slices (the cfg2 shape is 1 GB in fp64)
        ref = wav[n0:n0 + 8].unfold(1, win, hop).double() @ w.double().t()
        ref = torch.relu(ref) if relu else ref
        check(y[n0:n0 + 8], ref)
        check(y2[n0:n0 + 8], ref, 1e-5)


@pytest.mark.parametrize("with_res", [True, False])
def test_tc_layernorm_epilogue(ops, with_res):
    """Linear -> nn.LayerNorm -> + residual of the DPRNN blocks (dprnn.py:161-163,173-175) in the GEMM epilogue: the 128
    channels of a frame are the 128 TMEM lanes of the leader CTA.  Ragged last tile; SIMT back end = GEMM + row-norm."""
    B, Rr, M, K = 2, 333, 128, 256
    x, w, bias = rnd(B, Rr, K, seed=1, scale=2), rnd(M, K, seed=2, scale=0.1), rnd(M, seed=3)
    g, bt, res = rnd(M, seed=4) + 1.5, rnd(M, seed=5), rnd(B, Rr, M, seed=6)
    pk = ops.pack_weights(w, M, K, K)
    kw = dict(bias=bias, ln=(g, bt, 1e-5), residual=res if with_res else None)
    y, _ = ops.linear(x, w, w_packed=pk, backend=ops.GEMM_TCGEN05, **kw)
    lin = x.double() @ w.double().t() + bias.double()
    ref = F.layer_norm(lin, (M,), g.double(), bt.double(), 1e-5) + (res.double() if with_res else 0)
    check(y, ref, 1e-4)  # LayerNorm divides by the row's std (~0.5 here): the GEMM's 5e-5 grows accordingly
    y2, _ = ops.linear(x, w, backend=ops.GEMM_SIMT, **kw)
    check(y2, ref, 1e-5)


def test_tc_fused_prologue_epilogue_stats(ops):
    B, Rr, M, K = 3, 413, 512, 512
    x, w = rnd(B, Rr, K, seed=1, scale=3), rnd(M, K, seed=2, scale=0.05)
    sc, sh, slope = rnd(B, K, seed=3) + 1.5, rnd(B, K, seed=4), torch.tensor([0.2], device=DEV)
    bias, bb, res = rnd(M, seed=5), rnd(B, M, seed=6), rnd(B, Rr, M, seed=7)
    pk = ops.pack_weights(w, M, K, K)
    pro = ops.Prologue(ops.PRO_AFFINE, ops.ACT_PRELU, sc, sh, K, None, slope)
    y, part = ops.linear(x, w, pro=pro, bias=bias, bias_batch=bb, residual=res, want_stats=True, w_packed=pk, backend=ops.GEMM_TCGEN05)
    xin = F.prelu((x * sc.unsqueeze(1) + sh.unsqueeze(1)), slope).double()
    ref = xin @ w.double().t() + bias.double() + bb.double().unsqueeze(1) + res.double()
    check(y, ref)
    # Welford partials of the tile outputs -> same folded affine as torch's mean / biased var
    scale, shift = ops.stats_finalize(part, None, None, 1e-8, M)
    mu = ref.mean(dim=(1, 2))
    rstd = 1 / torch.sqrt(ref.var(dim=(1, 2), unbiased=False) + 1e-8)
    assert (scale[:, 0].double() - rstd).abs().max() <= 1e-5 * rstd.abs().max()
    assert (shift[:, 0].double() + mu * rstd).abs().max() <= 5e-5
    # and the two back ends agree with each other
    y2, _ = ops.linear(x, w, pro=pro, bias=bias, bias_batch=bb, residual=res, backend=ops.GEMM_SIMT)
    check(y, y2.double())
    # fused finalize: the CTA that completes an item's last tile emits the same folded affine, bit for bit, and leaves
    # the per-item counters at zero (run twice: the second launch depends on it)
    gamma, beta = rnd(M, seed=8) + 1.5, rnd(M, seed=9)
    sc_ref, sh_ref = ops.stats_finalize(part, gamma, beta, 1e-8, M)
    for _ in range(2):
        _, fa = ops.linear(x, w, pro=pro, bias=bias, bias_batch=bb, residual=res, want_stats=True, fin=(gamma, beta, 1e-8),
                           w_packed=pk, backend=ops.GEMM_TCGEN05)
        assert torch.equal(fa.scale, sc_ref) and torch.equal(fa.shift, sh_ref)
    # the exact-fp32 back end has no fused finalize: the library runs the merge as a follow-up launch, same interface
    _, fa2 = ops.linear(x, w, pro=pro, bias=bias, bias_batch=bb, residual=res, want_stats=True, fin=(gamma, beta, 1e-8),
                        backend=ops.GEMM_SIMT)
    assert (fa2.scale - sc_ref).abs().max() <= 1e-5 * sc_ref.abs().max()


def test_tc_weight_columns_view_and_relu(ops):
    """in_conv of a conditioned block: the packed image is built from the first C columns of a [H, C+E] weight."""
    H, Cc, E = 256, 512, 192
    w_full, x = rnd(H, Cc + E, seed=1, scale=0.05), rnd(2, 200, Cc, seed=2)
    pk = ops.pack_weights(w_full, H, Cc, Cc + E)
    y, _ = ops.linear(x, w_full, K=Cc, w_row_stride=Cc + E, epi_act=ops.ACT_RELU, w_packed=pk, backend=ops.GEMM_TCGEN05)
    check(y, torch.relu(x.double() @ w_full[:, :Cc].double().t()))


def test_tc_many_tiles_persistent(ops):
    """more tiles than SMs: every CTA walks several tiles through both TMEM accumulator buffers"""
    B, Rr, M, K = 8, 3999, 512, 128
    x, w = rnd(B, Rr, K, seed=1), rnd(M, K, seed=2, scale=0.1)
    pk = ops.pack_weights(w, M, K, K)
    y, _ = ops.linear(x, w, w_packed=pk, backend=ops.GEMM_TCGEN05)
    check(y, x.double() @ w.double().t())


def test_tc_nan_inf_rows_stay_local(ops):
    x, w = rnd(1, 256, 64, seed=1), rnd(256, 64, seed=2)
    x[0, 100:, :] = float("inf")
    y, _ = ops.linear(x, w, w_packed=ops.pack_weights(w, 256, 64, 64), backend=ops.GEMM_TCGEN05)
    assert torch.isfinite(y[0, :100]).all() and not torch.isfinite(y[0, 100:]).any()


def test_tc_ineligible_shapes_are_refused(ops):
    x, w = rnd(1, 10, 60, seed=1), rnd(130, 60, seed=2)
    assert ops.pack_weights(w, 130, 60, 60) is None
    assert ops.pack_weights(rnd(130, 64, seed=3), 130, 64, 64) is None  # channels must come in multiples of 32
    with pytest.raises(NotImplementedError):
        ops.linear(x, w, backend=ops.GEMM_TCGEN05)


@pytest.mark.parametrize("B,Rr,M,K", [(2, 300, 256, 128), (1, 747, 512, 128), (3, 130, 128, 64)])
def test_tc_affine_tanh_prologue(ops, B, Rr, M, K):
    """Second conv of AttentiveStatisticsPooling (lobe/pooling.py:71-86,104-105): eval-BatchNorm folded to a per-channel affine,
    then tanh, applied to the operand on load (batch-independent scale / shift)."""
    x, w, bias = rnd(B, Rr, K, seed=1, scale=2), rnd(M, K, seed=2, scale=0.1), rnd(M, seed=3)
    sc, sh = rnd(K, seed=4) + 1.5, rnd(K, seed=5)
    pk = ops.pack_weights(w, M, K, K)
    assert pk is not None
    pro = ops.Prologue(ops.PRO_AFFINE, ops.ACT_TANH, sc, sh, 0)
    y, _ = ops.linear(x, w, pro=pro, bias=bias, w_packed=pk, backend=ops.GEMM_TCGEN05)
    ref = torch.tanh(x.double() * sc.double() + sh.double()) @ w.double().t() + bias.double()
    check(y, ref)
    y2, _ = ops.linear(x, w, pro=pro, bias=bias, backend=ops.GEMM_SIMT)
    check(y2, ref, 1e-5)


@pytest.mark.parametrize("B,Rr,M,K", [(2, 10000, 512, 512), (40, 517, 512, 64), (3, 7000, 512, 192), (64, 300, 1024, 128), (1, 19000, 512, 256)])
def test_tc_wide_tiles(ops, B, Rr, M, K):
    """gemm_wide_kernel (256-frame tiles, MMA N = 256, per-block accumulator hand-over with the skewed MMA order at the tile
    boundaries): chosen for M % 512 == 0 once there is a full wave of tiles.  K = 64 (fewer stages than the skew holds), 192,
    128 (exactly the ring), 256, 512; ragged last tiles; M = 1024 (two channel groups); affine + PReLU prologue, bias,
    per-item bias, residual, Welford partials - against fp64 and the exact-fp32 back end."""
    x, w = rnd(B, Rr, K, seed=1, scale=3), rnd(M, K, seed=2, scale=0.05)
    sc, sh, slope = rnd(B, K, seed=3) + 1.5, rnd(B, K, seed=4), torch.tensor([0.2], device=DEV)
    bias, bb, res = rnd(M, seed=5), rnd(B, M, seed=6), rnd(B, Rr, M, seed=7)
    pk = ops.pack_weights(w, M, K, K)
    pro = ops.Prologue(ops.PRO_AFFINE, ops.ACT_PRELU, sc, sh, K, None, slope)
    ops.path_log = []
    y, part = ops.linear(x, w, pro=pro, bias=bias, bias_batch=bb, residual=res, want_stats=True, w_packed=pk, backend=ops.GEMM_TCGEN05)
    y0, _ = ops.linear(x, w, w_packed=pk, backend=ops.GEMM_TCGEN05)  # no prologue, no epilogue extras, no statistics
    paths, ops.path_log = ops.path_log, None
    # the call without a residual has at least one 256-frame tile per CTA pair -> gemm_wide_kernel; with a residual, launches
    # below two tiles per pair go to one-block tiles with 64-k stages (gemm_pair_few_tiles), the largest shape stays wide
    assert paths[1][1] == 3 and paths[0][1] == (3 if B * ((Rr + 255) // 256) * (M // 512) >= 148 else 2), paths
    for b0 in range(0, B, 8):
        sl = slice(b0, b0 + 8)
        xin = F.prelu((x[sl] * sc[sl].unsqueeze(1) + sh[sl].unsqueeze(1)), slope).double()
        ref = xin @ w.double().t() + bias.double() + bb[sl].double().unsqueeze(1) + res[sl].double()
        check(y[sl], ref)
        check(y0[sl], x[sl].double() @ w.double().t())
    scale, shift = ops.stats_finalize(part, None, None, 1e-8, M)
    yd = y.double()
    mu, rstd = yd.mean(dim=(1, 2)), 1 / torch.sqrt(yd.var(dim=(1, 2), unbiased=False) + 1e-8)
    assert (scale[:, 0].double() - rstd).abs().max() <= 1e-5 * rstd.abs().max()
    assert (shift[:, 0].double() + mu * rstd).abs().max() <= 5e-5
    y2, _ = ops.linear(x, w, pro=pro, bias=bias, bias_batch=bb, residual=res, backend=ops.GEMM_SIMT)
    check(y, y2.double())
    # out_conv form (bias + residual, no statistics): full tiles take the epilogue that requests the residual one chunk ahead
    y4, _ = ops.linear(x, w, pro=pro, bias=bias, residual=res, w_packed=pk, backend=ops.GEMM_TCGEN05)
    y5, _ = ops.linear(x, w, pro=pro, bias=bias, residual=res, backend=ops.GEMM_SIMT)
    check(y4, y5.double())


@pytest.mark.parametrize("B,Rr,M,K", [(1, 3999, 512, 512), (3, 497, 512, 256), (1, 640, 1024, 64), (2, 900, 384, 128)])
def test_tc_few_tiles_one_block_from_two_block_image(ops, B, Rr, M, K):
    """512-channel layers with fewer than one 256-frame tile per CTA pair (batch 1, the TSE model's 497-frame items): one
    256-channel block per tile with 64-k stages, the hi / lo parts fetched out of the two-block packed image (w_from2);
    M = 384 pads its second block.  Prologue, bias, per-item bias, residual, statistics and the fused finalize."""
    x, w = rnd(B, Rr, K, seed=1, scale=3), rnd(M, K, seed=2, scale=0.05)
    sc, sh, slope = rnd(B, K, seed=3) + 1.5, rnd(B, K, seed=4), torch.tensor([0.2], device=DEV)
    bias, bb, res = rnd(M, seed=5), rnd(B, M, seed=6), rnd(B, Rr, M, seed=7)
    g, bt = rnd(M, seed=8) + 1.5, rnd(M, seed=9)
    pk = ops.pack_weights(w, M, K, K)
    pro = ops.Prologue(ops.PRO_AFFINE, ops.ACT_PRELU, sc, sh, K, None, slope)
    ops.path_log = []
    y, fold = ops.linear(x, w, pro=pro, bias=bias, bias_batch=bb, residual=res, want_stats=True, fin=(g, bt, 1e-8), w_packed=pk,
                         backend=ops.GEMM_TCGEN05)
    paths, ops.path_log = ops.path_log, None
    assert [p for _, p in paths] == [2], paths
    ref = F.prelu((x * sc.unsqueeze(1) + sh.unsqueeze(1)), slope).double() @ w.double().t() + bias.double() + bb.double().unsqueeze(1) + res.double()
    check(y, ref)
    rstd = 1 / torch.sqrt(ref.var(dim=(1, 2), unbiased=False) + 1e-8)
    assert (fold.scale.double() - g.double() * rstd.unsqueeze(1)).abs().max() <= 1e-5 * (g.abs().max() * rstd.max()).item()
    mu = ref.mean(dim=(1, 2))
    assert (fold.shift.double() - (bt.double() - (mu * rstd).unsqueeze(1) * g.double())).abs().max() <= 1e-4


def test_tc_wide_slopes_and_nonfinite(ops):
    """PReLU slopes outside (0, 1] take the select instead of max(u, slope*u); Inf / NaN inputs stay in their rows and come
    out non-finite (the reference's look-ahead probe, base_nn.py:740-777, depends on it)."""
    B, Rr, M, K = 2, 9600, 512, 128
    x, w = rnd(B, Rr, K, seed=1, scale=2), rnd(M, K, seed=2, scale=0.1)
    sc, sh = rnd(B, K, seed=3) + 1.5, rnd(B, K, seed=4)
    pk = ops.pack_weights(w, M, K, K)
    for sl in (0.0, 1.0, 1.7, -0.3):
        slope = torch.tensor([sl], device=DEV)
        pro = ops.Prologue(ops.PRO_AFFINE, ops.ACT_PRELU, sc, sh, K, None, slope)
        y, _ = ops.linear(x, w, pro=pro, w_packed=pk, backend=ops.GEMM_TCGEN05)
        check(y, F.prelu(x * sc.unsqueeze(1) + sh.unsqueeze(1), slope).double() @ w.double().t())
    x[0, 5000:, :] = float("inf")
    x[1, :100, 3] = float("nan")
    pro = ops.Prologue(ops.PRO_AFFINE, ops.ACT_PRELU, sc.abs() + 0.1, sh, K, None, torch.tensor([0.25], device=DEV))
    y, _ = ops.linear(x, w, pro=pro, w_packed=pk, backend=ops.GEMM_TCGEN05)
    assert torch.isfinite(y[0, :5000]).all() and not torch.isfinite(y[0, 5000:]).any()
    assert not torch.isfinite(y[1, :100]).any() and torch.isfinite(y[1, 100:]).all()


@pytest.mark.parametrize("B,Rr,M,K", [(2, 333, 128, 256), (1, 5000, 32, 512), (3, 129, 64, 512), (5, 77, 96, 64), (1, 40000, 128, 128),
                                       (64, 300, 64, 192)])
def test_tc_rows_kernel_few_channels(ops, B, Rr, M, K):
    """gemm_rows_kernel (ps_gemm_rows.cu): M <= 128 output channels on the MMA's N side, frames on the 128 TMEM lanes, the
    packed weight matrix resident in shared memory.  BN = 32 / 64 / 128 (M = 96 pads to 128), one and many tiles per CTA,
    ragged last tiles, folded norm affine + PReLU prologue, bias, per-item bias, PReLU epilogue, residual, Welford partials,
    the fused-finalize request (served by a follow-up merge) - against fp64 and the exact-fp32 back end."""
    x, w = rnd(B, Rr, K, seed=1, scale=3), rnd(M, K, seed=2, scale=0.05)
    sc, sh, slope = rnd(B, K, seed=3) + 1.5, rnd(B, K, seed=4), torch.tensor([0.2], device=DEV)
    bias, bb, res = rnd(M, seed=5), rnd(B, M, seed=6), rnd(B, Rr, M, seed=7)
    pk = ops.pack_weights(w, M, K, K)
    assert pk is not None and pk.numel() == 256 * K * 4 + (32 if M <= 32 else 64 if M <= 64 else 128) * K * 4
    pro = ops.Prologue(ops.PRO_AFFINE, ops.ACT_PRELU, sc, sh, K, None, slope)
    ops.path_log = []
    y, part = ops.linear(x, w, pro=pro, bias=bias, bias_batch=bb, residual=res, want_stats=True, w_packed=pk, backend=ops.GEMM_TCGEN05)
    y0, _ = ops.linear(x, w, w_packed=pk, backend=ops.GEMM_TCGEN05, epi_act=ops.ACT_PRELU, epi_slope=slope)
    paths, ops.path_log = ops.path_log, None
    assert [p for _, p in paths] == [4, 4], paths
    xin = F.prelu((x * sc.unsqueeze(1) + sh.unsqueeze(1)), slope).double()
    ref = xin @ w.double().t() + bias.double() + bb.double().unsqueeze(1) + res.double()
    check(y, ref)
    check(y0, F.prelu(x.double() @ w.double().t(), slope.double()))
    scale, shift = ops.stats_finalize(part, None, None, 1e-8, M)
    mu, rstd = ref.mean(dim=(1, 2)), 1 / torch.sqrt(ref.var(dim=(1, 2), unbiased=False) + 1e-8)
    assert (scale[:, 0].double() - rstd).abs().max() <= 1e-5 * rstd.abs().max()
    assert (shift[:, 0].double() + mu * rstd).abs().max() <= 5e-5
    g, bt = rnd(M, seed=8) + 1.5, rnd(M, seed=9)
    y3, fold = ops.linear(x, w, bias=bias, want_stats=True, fin=(g, bt, 1e-8), w_packed=pk, backend=ops.GEMM_TCGEN05)
    ref3 = x.double() @ w.double().t() + bias.double()
    check(y3, ref3)
    rstd3 = 1 / torch.sqrt(ref3.var(dim=(1, 2), unbiased=False) + 1e-8)
    assert (fold.scale.double() - g.double() * rstd3.unsqueeze(1)).abs().max() <= 1e-5 * (g.abs().max() * rstd3.max()).item()
    y2, _ = ops.linear(x, w, pro=pro, bias=bias, bias_batch=bb, residual=res, backend=ops.GEMM_SIMT)
    check(y, y2.double())


@pytest.mark.parametrize("M,K,with_res", [(128, 256, True), (128, 256, False), (64, 128, True), (96, 192, True), (32, 1024, False)])
def test_tc_rows_kernel_layernorm(ops, M, K, with_res):
    """Linear -> nn.LayerNorm -> + residual (dprnn.py:161-163,173-175) on the few-channel kernel: an epilogue thread owns one
    frame (a TMEM lane), so mean and variance over the M channels are thread-local.  In place over the residual as well."""
    B, Rr = 3, 1333
    x, w, bias = rnd(B, Rr, K, seed=1, scale=2), rnd(M, K, seed=2, scale=0.1), rnd(M, seed=3)
    g, bt, res = rnd(M, seed=4) + 1.5, rnd(M, seed=5), rnd(B, Rr, M, seed=6)
    pk = ops.pack_weights(w, M, K, K)
    kw = dict(bias=bias, ln=(g, bt, 1e-5), residual=res if with_res else None)
    ops.path_log = []
    y, _ = ops.linear(x, w, w_packed=pk, backend=ops.GEMM_TCGEN05, **kw)
    paths, ops.path_log = ops.path_log, None
    assert [p for _, p in paths] == [4], paths
    lin = x.double() @ w.double().t() + bias.double()
    ref = F.layer_norm(lin, (M,), g.double(), bt.double(), 1e-5) + (res.double() if with_res else 0)
    check(y, ref, 1e-4)
    if with_res:  # out aliasing the residual (a thread reads exactly the elements it then writes)
        buf = res.clone()
        ops.linear(x, w, w_packed=pk, backend=ops.GEMM_TCGEN05, out=buf, **dict(kw, residual=buf))
        check(buf, ref, 1e-4)


@pytest.mark.parametrize("act", ["none", "relu", "sigmoid"])
def test_tc_rows_kernel_mask_prologue_and_views(ops, act):
    """The decoder GEMM (mask apply on load, M = 32) at many tiles per CTA; overlapping operand rows (row stride < K: the
    U-Net shell's tap windows) and a strided output view; NaN / Inf rows stay local."""
    B, Rr, M, K = 2, 30000, 32, 512
    x, mk, w = rnd(B, Rr, K, seed=1, scale=2), rnd(B, Rr, K, seed=2, scale=2), rnd(M, K, seed=3, scale=0.1)
    code = {"none": ops.ACT_NONE, "relu": ops.ACT_RELU, "sigmoid": ops.ACT_SIGMOID}[act]
    f = {"none": lambda t: t, "relu": torch.relu, "sigmoid": torch.sigmoid}[act]
    pk = ops.pack_weights(w, M, K, K)
    ops.path_log = []
    y, _ = ops.linear(x, w, pro=ops.Prologue(ops.PRO_MASK, code, x2=mk), w_packed=pk, backend=ops.GEMM_TCGEN05)
    paths, ops.path_log = ops.path_log, None
    assert [p for _, p in paths] == [4], paths
    check(y, (x.double() * f(mk.double())) @ w.double().t())
    if act != "none":
        return
    N, L, win, hop, Mo = 2, 16000, 256, 64, 64
    wav, w2 = rnd(N, L, seed=4), rnd(Mo, win, seed=5, scale=0.05)
    T = (L - win) // hop + 1
    pk2 = ops.pack_weights(w2, Mo, win, win)
    big = torch.zeros(N, T, 3 * Mo, device=DEV)
    ops.gemm(wav, w2, batch=N, rows=T, M=Mo, K=win, x_batch_stride=L, x_row_stride=hop, w_row_stride=win, w_packed=pk2,
             backend=ops.GEMM_TCGEN05, out=big[:, :, Mo:], y_strides=(T * 3 * Mo, 3 * Mo))
    check(big[:, :, Mo:2 * Mo], wav.unfold(1, win, hop).double() @ w2.double().t())
    assert big[:, :, :Mo].abs().max() == 0 and big[:, :, 2 * Mo:].abs().max() == 0
    xn = rnd(1, 1000, 128, seed=6)
    xn[0, 500:, :] = float("inf")
    xn[0, :10, 7] = float("nan")
    w3 = rnd(128, 128, seed=7, scale=0.1)
    yn, _ = ops.linear(xn, w3, w_packed=ops.pack_weights(w3, 128, 128, 128), backend=ops.GEMM_TCGEN05)
    assert not torch.isfinite(yn[0, :10]).any() and torch.isfinite(yn[0, 10:500]).all() and not torch.isfinite(yn[0, 500:]).any()


@pytest.mark.parametrize("B,Rr,M,K", [(2, 3000, 128, 512), (1, 9000, 64, 1536), (3, 700, 32, 2048), (1, 20000, 128, 768), (2, 333, 96, 1024)])
def test_tc_rows_kernel_streamed_weights(ops, B, Rr, M, K):
    """gemm_rows_kernel with a packed weight matrix of more than 128 KB (the U-Net shell's tap windows, out_conv of the
    128-channel Conv-TasNet reading): the 32-k weight blocks travel through the six-stage ring with the operand.  Plain,
    affine + PReLU (K <= 1024), mask prologue, bias / per-item bias / residual / statistics, LayerNorm epilogue."""
    x, w = rnd(B, Rr, K, seed=1, scale=2), rnd(M, K, seed=2, scale=0.03)
    bias, bb, res = rnd(M, seed=5), rnd(B, M, seed=6), rnd(B, Rr, M, seed=7)
    pk = ops.pack_weights(w, M, K, K)
    ops.path_log = []
    y, part = ops.linear(x, w, bias=bias, bias_batch=bb, residual=res, want_stats=True, w_packed=pk, backend=ops.GEMM_TCGEN05)
    mk = rnd(B, Rr, K, seed=8, scale=2)
    ym, _ = ops.linear(x, w, pro=ops.Prologue(ops.PRO_MASK, ops.ACT_SIGMOID, x2=mk), w_packed=pk, backend=ops.GEMM_TCGEN05)
    g, bt = rnd(M, seed=9) + 1.5, rnd(M, seed=10)
    yl, _ = ops.linear(x, w, bias=bias, ln=(g, bt, 1e-5), residual=res, w_packed=pk, backend=ops.GEMM_TCGEN05)
    paths, ops.path_log = ops.path_log, None
    assert [p for _, p in paths] == [4, 4, 4], paths
    ref = x.double() @ w.double().t() + bias.double() + bb.double().unsqueeze(1) + res.double()
    check(y, ref)
    scale, shift = ops.stats_finalize(part, None, None, 1e-8, M)
    rstd = 1 / torch.sqrt(ref.var(dim=(1, 2), unbiased=False) + 1e-8)
    assert (scale[:, 0].double() - rstd).abs().max() <= 1e-5 * rstd.abs().max()
    check(ym, (x.double() * torch.sigmoid(mk.double())) @ w.double().t())
    lin = x.double() @ w.double().t() + bias.double()
    check(yl, F.layer_norm(lin, (M,), g.double(), bt.double(), 1e-5) + res.double(), 2e-4)
    if K <= 1024:
        sc, sh, slope = rnd(B, K, seed=3) + 1.5, rnd(B, K, seed=4), torch.tensor([0.2], device=DEV)
        ops.path_log = []
        ya, _ = ops.linear(x, w, pro=ops.Prologue(ops.PRO_AFFINE, ops.ACT_PRELU, sc, sh, K, None, slope), w_packed=pk, backend=ops.GEMM_TCGEN05)
        paths, ops.path_log = ops.path_log, None
        assert [p for _, p in paths] == [4], paths
        check(ya, F.prelu(x * sc.unsqueeze(1) + sh.unsqueeze(1), slope).double() @ w.double().t())


@pytest.mark.parametrize("B,Rr,M,K", [(3, 333, 96, 128), (2, 130, 32, 512), (2, 257, 128, 768), (1, 3999, 512, 512), (3, 497, 384, 128)])
def test_tc_outputs_stay_inside_their_view(ops, B, Rr, M, K):
    """Guard bands around a strided output view (compute-sanitizer is not available on the GPU pool): ragged last tiles, channel
    counts below the MMA width (96 -> 128, 384 -> 512) and the few-tile path must not write a byte outside [B, rows, M] - rows
    past the end of an item are computed (the operand re-reads the last row) but never stored."""
    x, w, bias = rnd(B, Rr, K, seed=1, scale=2), rnd(M, K, seed=2, scale=0.05), rnd(M, seed=3)
    pk = ops.pack_weights(w, M, K, K)
    sentinel = 12345.0
    big = torch.full((B, Rr + 3, M + 32), sentinel, device=DEV)
    ops.gemm(x, w, batch=B, rows=Rr, M=M, K=K, x_batch_stride=Rr * K, x_row_stride=K, w_row_stride=K, bias=bias, w_packed=pk,
             backend=ops.GEMM_TCGEN05, out=big, y_strides=((Rr + 3) * (M + 32), M + 32))
    check(big[:, :Rr, :M], x.double() @ w.double().t() + bias.double())
    assert (big[:, Rr:, :] == sentinel).all() and (big[:, :, M:] == sentinel).all()
