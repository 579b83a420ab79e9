"""Thin host-side wrappers: torch CUDA tensors in, C-ABI calls out.

PyTorch is plumbing here (device memory, streams); every arithmetic step of the
separator forward is a kernel in ``libpuresound_b200.so``.  Activations are
frames-major ``[batch, rows, channels]`` fp32 contiguous.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (ACT_NONE, ACT_PRELU, ACT_RELU, ACT_SIGMOID, ACT_TANH, GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05,  # noqa: F401
                   PRO_AFFINE, PRO_MASK, PRO_NONE, PRO_ROWNORM)

Tensor = torch.Tensor

#: kernels launched through this module since the last reset (bench.py's gpu_launches)
launch_count = 0
#: when a list, every ps_gemm / ps_lstm / ps_dwconv launch appends (kind, start_event, end_event, shape) - bench.py's live
#: per-kernel roofline timing (CUDA events on the launching stream); shape = (rows, M, K) for "gemm",
#: (n_seq, L, H, D, K_in) for "lstm" (K_in > 0 when the input projection is fused), (batch, T, C) for "dwconv"; a fifth
#: element carries the launch's algorithmic bytes beyond the shape (residual / mask operand of a GEMM)
kernel_events = None


#: when a list, every ps_gemm call appends ((rows, M, K), path) with path = ps_gemm_path(): 0 exact-fp32 CUDA cores,
#: 1 single-CTA tcgen05, 2 CTA-pair tcgen05 (128-frame tiles), 3 wide CTA-pair tcgen05 (256-frame tiles), 4 few-channel tcgen05 (M <= 128)
path_log = None
GEMM_PATH_NAMES = {0: "gemm_simt_kernel (fp32 CUDA cores)", 1: "gemm_tc_kernel (tcgen05, one CTA)", 2: "gemm_pair_kernel (tcgen05 cta_group::2, 128-frame tiles)",
                   3: "gemm_wide_kernel (tcgen05 cta_group::2, 256-frame tiles)",
                   4: "gemm_rows_kernel (tcgen05, M <= 128 channels on the MMA's N side, weights resident)"}


class _Timed:
    """Context manager: bracket one launch with a CUDA-event pair when bench.py asked for per-kernel timing."""

    __slots__ = ("kind", "shape", "ev", "extra_bytes")

    def __init__(self, kind, shape, extra_bytes=0):
        # extra_bytes: algorithmic HBM bytes of this launch beyond what the shape implies (a GEMM's residual / mask operand)
        self.kind, self.shape, self.ev, self.extra_bytes = kind, shape, None, extra_bytes

    def __enter__(self):
        if kernel_events is not None:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()
        return self

    def __exit__(self, *a):
        if self.ev is not None:
            self.ev[1].record()
            kernel_events.append((self.kind, self.ev[0], self.ev[1], self.shape, self.extra_bytes))
#: force a GEMM back end for every call (tests / A-B runs); None = per-call choice
force_gemm_backend: Optional[int] = None


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t: Tensor, what: str) -> Tensor:
    if not (t.is_cuda and t.dtype == torch.float32):
        raise TypeError(f"{what}: expected a CUDA float32 tensor, got {t.device} {t.dtype} (no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError(f"{what}: expected a contiguous tensor")
    return t


def _dev(t: Tensor, what: str) -> Tensor:
    if not (t.is_cuda and t.dtype == torch.float32):
        raise TypeError(f"{what}: expected a CUDA float32 tensor, got {t.device} {t.dtype} (no CPU fallback)")
    return t


def _launched(n: int = 1):
    global launch_count
    launch_count += n


def require_device() -> None:
    """Fail loudly unless a B200-class device and the built library are present."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.EngineMissing("puresound_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    ok = lib.ps_device_ok()
    if ok != 1:
        raise _lib.EngineMissing(f"puresound_b200 kernels are built for sm_100a only (device check returned {ok})")


@dataclass
class Prologue:
    """What a consumer kernel applies to its input on load."""
    mode: int = PRO_NONE
    act: int = ACT_NONE
    a: Optional[Tensor] = None          # AFFINE scale [B or 1, K] / ROWNORM gamma [K]
    b: Optional[Tensor] = None
    batch_stride: int = 0
    rowstats: Optional[Tensor] = None   # [B, rows, 2]
    slope: Optional[Tensor] = None      # PReLU weight (1,)
    x2: Optional[Tensor] = None         # MASK operand


NO_PRO = Prologue()


@dataclass
class FoldedAffine:
    """gLN / gGN folded to the per-item per-channel affine a consumer prologue applies: scale, shift [B, C]."""
    scale: Tensor
    shift: Tensor
    counter: Optional[Tensor] = None    # the producer's per-item tile counter (kept alive with the result)


def _set_fin(d, fin, batch: int, Cn: int, device):
    """fin = (gamma, beta, eps): ask the producer kernel to also emit the folded norm affine (fused ps_stats_finalize).
    The per-item tile counter is allocated (zeroed) per call and kept alive by the returned FoldedAffine: no buffer is
    shared between launches, streams or captured graphs."""
    gamma, beta, eps = fin
    scale = torch.empty(batch, Cn, device=device, dtype=torch.float32)
    shift = torch.empty_like(scale)
    counter = torch.zeros(batch, device=device, dtype=torch.int32)
    d.fin_gamma, d.fin_beta, d.fin_eps = _p(gamma), _p(beta), float(eps)
    d.fin_scale, d.fin_shift = scale.data_ptr(), shift.data_ptr()
    d.fin_counter = counter.data_ptr()
    return FoldedAffine(scale, shift, counter)


def gemm(
    X: Tensor, W: Tensor, *, batch: int, rows: int, M: int, K: int,
    x_batch_stride: int, x_row_stride: int, w_row_stride: int,
    pro: Prologue = NO_PRO, bias: Optional[Tensor] = None, bias_batch: Optional[Tensor] = None,
    epi_act: int = ACT_NONE, epi_slope: Optional[Tensor] = None, residual: Optional[Tensor] = None,
    want_stats: bool = False, out: Optional[Tensor] = None, backend: int = GEMM_AUTO,
    w_packed: Optional[Tensor] = None, fin=None, ln=None,
    y_strides: Optional[tuple] = None, res_strides: Optional[tuple] = None,
):
    """Y[b,r,m] = epi(sum_k pro(X[b,r,k]) W[m,k]); returns (Y [batch, rows, M], stats partials or None) - or, with
    want_stats and fin=(gamma, beta, eps), (Y, FoldedAffine): the producer also finalizes the gLN/gGN statistics.
    ln=(weight, bias, eps): Y = residual + LayerNorm_M(X W^T + bias) (fused epilogue on the tcgen05 kernel when M == 128).
    y_strides / res_strides = (batch stride, row stride) in floats when `out` / `residual` are views into larger buffers
    (`out` is then the tensor whose data_ptr is the first output element)."""
    lib = _lib.load()
    _req(X, "gemm X")
    _dev(W, "gemm W")  # may be a row-strided view (embedding columns of in_conv)
    Y = out if out is not None else torch.empty(batch, rows, M, device=X.device, dtype=torch.float32)
    partials = None
    if want_stats:
        slots = lib.ps_gemm_stats_slots(rows, M)
        partials = torch.empty(batch, slots, 3, device=X.device, dtype=torch.float32)
    d = _lib.GemmDesc()
    d.batch, d.rows, d.M, d.K = batch, rows, M, K
    d.X, d.x_batch_stride, d.x_row_stride = X.data_ptr(), x_batch_stride, x_row_stride
    d.W, d.w_row_stride = W.data_ptr(), w_row_stride
    d.Y, (d.y_batch_stride, d.y_row_stride) = Y.data_ptr(), (y_strides if y_strides is not None else (rows * M, M))
    d.pro_mode, d.pro_act = pro.mode, pro.act
    d.pro_a, d.pro_b, d.pro_batch_stride = _p(pro.a), _p(pro.b), pro.batch_stride
    d.pro_rowstats, d.pro_slope, d.X2 = _p(pro.rowstats), _p(pro.slope), _p(pro.x2)
    d.bias, d.bias_batch = _p(bias), _p(bias_batch)
    d.epi_act = epi_act
    d.backend = force_gemm_backend if force_gemm_backend is not None else backend
    d.epi_slope = _p(epi_slope)
    if residual is not None:
        d.residual = residual.data_ptr()
        d.res_batch_stride, d.res_row_stride = res_strides if res_strides is not None else (rows * M, M)
    d.stats_partials = _p(partials)
    d.W_packed = _p(w_packed)
    folded = _set_fin(d, fin, batch, M, X.device) if (want_stats and fin is not None) else None
    if ln is not None:
        d.ln_gamma, d.ln_beta, d.ln_eps = _p(ln[0]), _p(ln[1]), float(ln[2])
    if path_log is not None:
        path_log.append(((batch * rows, M, K), int(lib.ps_gemm_path(C.byref(d)))))
    extra = 4 * batch * rows * ((M if residual is not None else 0) + (K if pro.x2 is not None else 0))
    with _Timed("gemm", (batch * rows, M, K), extra):
        _lib.check(lib.ps_gemm(C.byref(d), _stream()), "ps_gemm")
    _launched()
    return Y, (folded if folded is not None else partials)


def pack_weights(W: Tensor, M: int, K: int, w_row_stride: int) -> Optional[Tensor]:
    """Pre-split (bf16 hi/lo) and pre-swizzle W[:M, :K] into the tcgen05 kernel's shared-memory tile image.
    Returns None when the shape is not eligible for the tensor-core back end (M % 256, K % 64)."""
    lib = _lib.load()
    nbytes = lib.ps_gemm_packed_bytes(M, K)
    if nbytes == 0:
        return None
    out = torch.empty(nbytes, device=W.device, dtype=torch.uint8)
    _lib.check(lib.ps_gemm_pack_weights(_dev(W, "pack W").data_ptr(), w_row_stride, M, K, out.data_ptr(), _stream()),
               "ps_gemm_pack_weights")
    _launched()
    return out


def linear(x: Tensor, W: Tensor, K: Optional[int] = None, w_row_stride: Optional[int] = None, **kw):
    """GEMM over a contiguous frames-major activation x [B, R, K]; W [M, >=K] row-major."""
    B, R, Kx = x.shape
    K = Kx if K is None else K
    return gemm(x, W, batch=B, rows=R, M=W.shape[0], K=K, x_batch_stride=R * Kx, x_row_stride=Kx,
                w_row_stride=W.stride(0) if w_row_stride is None else w_row_stride, **kw)


def linear_ln_residual(x: Tensor, W: Tensor, bias: Optional[Tensor], gamma: Tensor, beta: Tensor, eps: float, residual: Tensor,
                       w_packed: Optional[Tensor] = None) -> Tensor:
    """residual + LayerNorm(x W^T + bias) for x [1, P, K], W [M, K] (dprnn.py:161-163,173-175; skim.py SegLSTM; dpcrn.py /
    dparn.py blocks) in one GEMM (LayerNorm in the epilogue of the tcgen05 kernel when M == 128)."""
    y, _ = linear(x, W, bias=bias, w_packed=w_packed, ln=(gamma, beta, eps), residual=residual)
    return y


def stats_finalize(partials: Tensor, gamma: Optional[Tensor], beta: Optional[Tensor], eps: float, Cn: int):
    """gLN/gGN: merge Welford partials -> folded per-item affine (scale, shift) [B, C]."""
    lib = _lib.load()
    B, slots, _ = partials.shape
    scale = torch.empty(B, Cn, device=partials.device, dtype=torch.float32)
    shift = torch.empty_like(scale)
    _lib.check(lib.ps_stats_finalize(partials.data_ptr(), B, slots, _p(gamma), _p(beta), eps, Cn, scale.data_ptr(),
                                     shift.data_ptr(), None, _stream()), "ps_stats_finalize")
    _launched()
    return scale, shift


def stats_region(x: Tensor, *, batch: int, mid: int, rows: int, C_: int, strides: tuple, slots: int = 64) -> Tensor:
    """Welford partials [batch, slots, 3] of the strided region x[n, m, r, :C] (strides = (batch, mid, row) in floats)."""
    lib = _lib.load()
    part = torch.empty(batch, slots, 3, device=x.device, dtype=torch.float32)
    _lib.check(lib.ps_stats_region(_dev(x, "stats_region").data_ptr(), batch, mid, rows, C_, strides[0], strides[1], strides[2], slots,
                                   part.data_ptr(), _stream()), "ps_stats_region")
    _launched()
    return part


def bn_fold(weight, bias, running_mean, running_var, eps: float):
    lib = _lib.load()
    Cn = running_mean.numel()
    scale = torch.empty(Cn, device=running_mean.device, dtype=torch.float32)
    shift = torch.empty_like(scale)
    _lib.check(lib.ps_bn_fold(_p(weight), _p(bias), running_mean.data_ptr(), running_var.data_ptr(), eps, Cn,
                              scale.data_ptr(), shift.data_ptr(), _stream()), "ps_bn_fold")
    _launched()
    return scale, shift


def rowstats(x: Tensor, eps: float) -> Tensor:
    """x [..., C] contiguous -> [..., 2] (mean, rstd) per row."""
    lib = _lib.load()
    Cn = x.shape[-1]
    rows = x.numel() // Cn
    out = torch.empty(*x.shape[:-1], 2, device=x.device, dtype=torch.float32)
    _lib.check(lib.ps_rowstats(_req(x, "rowstats").data_ptr(), rows, Cn, Cn, eps, out.data_ptr(), _stream()), "ps_rowstats")
    _launched()
    return out


def dwconv(x: Tensor, w: Tensor, bias: Optional[Tensor], P: int, dilation: int, causal: bool, pro: Prologue = NO_PRO,
           want_stats: bool = False, fin=None):
    """x [B, T, C] -> (y [B, T, C], partials) - or (y, FoldedAffine) with want_stats and fin=(gamma, beta, eps)."""
    lib = _lib.load()
    B, T, Cn = _req(x, "dwconv x").shape
    y = torch.empty_like(x)
    partials = None
    if want_stats:
        partials = torch.empty(B, lib.ps_dwconv_stats_slots(T, Cn), 3, device=x.device, dtype=torch.float32)
    d = _lib.DwconvDesc()
    d.batch, d.T, d.C, d.P, d.dilation, d.causal = B, T, Cn, P, dilation, int(causal)
    d.x, d.y, d.w, d.bias = x.data_ptr(), y.data_ptr(), w.data_ptr(), _p(bias)
    d.pro_mode, d.pro_act = pro.mode, pro.act
    d.pro_a, d.pro_b, d.pro_batch_stride = _p(pro.a), _p(pro.b), pro.batch_stride
    d.pro_rowstats, d.pro_slope = _p(pro.rowstats), _p(pro.slope)
    d.stats_partials = _p(partials)
    folded = _set_fin(d, fin, B, Cn, x.device) if (want_stats and fin is not None) else None
    with _Timed("dwconv", (B, T, Cn)):
        _lib.check(lib.ps_dwconv(C.byref(d), _stream()), "ps_dwconv")
    _launched()
    return y, (folded if folded is not None else partials)


def rownorm(x: Tensor, w: Optional[Tensor], b: Optional[Tensor], eps: float, res: Optional[Tensor] = None,
            act: int = ACT_NONE, slope: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    lib = _lib.load()
    Cn = x.shape[-1]
    rows = x.numel() // Cn
    y = out if out is not None else torch.empty_like(x)
    _lib.check(lib.ps_rownorm(_req(x, "rownorm").data_ptr(), _p(res), y.data_ptr(), rows, Cn, _p(w), _p(b), eps, act,
                              _p(slope), _stream()), "ps_rownorm")
    _launched()
    return y


def ola(frames: Tensor, hop: int, wsum: Optional[Tensor], constraint: int) -> Tensor:
    """frames [B, T, win] -> wav [B, (T-1)*hop + win]; constraint 0 none / 1 clamp / 2 sigmoid."""
    lib = _lib.load()
    B, T, win = _req(frames, "ola").shape
    y = torch.empty(B, (T - 1) * hop + win, device=frames.device, dtype=torch.float32)
    _lib.check(lib.ps_ola(frames.data_ptr(), B, T, win, hop, _p(wsum), constraint, y.data_ptr(), _stream()), "ps_ola")
    _launched()
    return y


def mask_apply(feats: Tensor, mask: Tensor, act: int, is_complex: bool) -> Tensor:
    lib = _lib.load()
    Cn = feats.shape[-1]
    y = torch.empty_like(feats)
    _lib.check(lib.ps_mask_apply(_req(feats, "mask feats").data_ptr(), _req(mask, "mask").data_ptr(), y.data_ptr(),
                                 feats.numel() // Cn, Cn, act, int(is_complex), _stream()), "ps_mask_apply")
    _launched()
    return y


def magnitude(x: Tensor, drop_first: bool, log1p: bool) -> Tensor:
    lib = _lib.load()
    F2 = x.shape[-1]
    F = F2 // 2
    y = torch.empty(*x.shape[:-1], F - int(drop_first), device=x.device, dtype=torch.float32)
    _lib.check(lib.ps_magnitude(_req(x, "magnitude").data_ptr(), y.data_ptr(), x.numel() // F2, F, int(drop_first),
                                int(log1p), _stream()), "ps_magnitude")
    _launched()
    return y


def band_fill(x: Tensor, T: int, Cn: int, bounds: Tensor, value: float) -> Tensor:
    """In place: x[..., t, c] = value for c in [bounds[0], bounds[1]) or t in [bounds[2], bounds[3]); bounds = int32[4] on the device."""
    lib = _lib.load()
    if not (bounds.is_cuda and bounds.dtype == torch.int32 and bounds.numel() == 4):
        raise TypeError("band_fill: bounds must be 4 int32 on the device")
    _lib.check(lib.ps_band_fill(_req(x, "band_fill x").data_ptr(), x.numel() // (T * Cn), T, Cn, bounds.data_ptr(), float(value),
                                _stream()), "ps_band_fill")
    _launched()
    return x


def asp_pool(x: Tensor, logits: Tensor) -> Tensor:
    lib = _lib.load()
    B, T, Cn = _req(x, "asp x").shape
    out = torch.empty(B, 2 * Cn, device=x.device, dtype=torch.float32)
    _lib.check(lib.ps_asp_pool(x.data_ptr(), _req(logits, "asp logits").data_ptr(), B, T, Cn, out.data_ptr(), _stream()),
               "ps_asp_pool")
    _launched()
    return out


def l2normalize(x: Tensor) -> Tensor:
    lib = _lib.load()
    rows, E = _req(x, "l2normalize").shape
    y = torch.empty_like(x)
    _lib.check(lib.ps_l2normalize(x.data_ptr(), y.data_ptr(), rows, E, _stream()), "ps_l2normalize")
    _launched()
    return y


def segment(x: Tensor, K: int, S: int, overlap: bool) -> Tensor:
    lib = _lib.load()
    B, T, Cn = _req(x, "segment").shape
    seg = torch.empty(B, S, K, Cn, device=x.device, dtype=torch.float32)
    _lib.check(lib.ps_segment(x.data_ptr(), seg.data_ptr(), B, T, Cn, K, S, int(overlap), _stream()), "ps_segment")
    _launched()
    return seg


def merge(seg: Tensor, T: int, overlap: bool) -> Tensor:
    lib = _lib.load()
    B, S, K, Cn = _req(seg, "merge").shape
    y = torch.empty(B, T, Cn, device=seg.device, dtype=torch.float32)
    _lib.check(lib.ps_merge(seg.data_ptr(), y.data_ptr(), B, T, Cn, K, S, int(overlap), _stream()), "ps_merge")
    _launched()
    return y


def lstm(gx: Tensor, w_hh_t: Tensor, *, n_seq: int, L: int, H: int, D: int, inner: int, outer_stride: int,
         inner_stride: int, step_stride: int, h0: Optional[Tensor] = None, c0: Optional[Tensor] = None,
         want_state: bool = False, w_packed: Optional[Tensor] = None, gx_interleaved: bool = False):
    """gx [positions, D*4H] -> out [positions, D*H] (+ (hn, cn) [D, n_seq, H])."""
    lib = _lib.load()
    positions = gx.shape[0]
    out = torch.empty(positions, D * H, device=gx.device, dtype=torch.float32)
    hn = cn = None
    if want_state:
        hn = torch.empty(D, n_seq, H, device=gx.device, dtype=torch.float32)
        cn = torch.empty_like(hn)
    d = _lib.LstmDesc()
    d.n_seq, d.L, d.H, d.D = n_seq, L, H, D
    d.inner, d.outer_stride, d.inner_stride, d.step_stride = inner, outer_stride, inner_stride, step_stride
    d.gx, d.w_hh_t, d.h0, d.c0 = _req(gx, "lstm gx").data_ptr(), w_hh_t.data_ptr(), _p(h0), _p(c0)
    d.out, d.hn, d.cn = out.data_ptr(), _p(hn), _p(cn)
    d.w_packed = _p(w_packed)
    d.gx_interleaved = 1 if (gx_interleaved and w_packed is not None) else 0
    with _Timed("lstm", (n_seq, L, H, D, 0)):
        _lib.check(lib.ps_lstm(C.byref(d), _stream()), "ps_lstm")
    _launched()
    return out, ((hn, cn) if want_state else None)


def lstm_pack_weights(w_hh_t: Tensor, H: int, D: int) -> Optional[Tensor]:
    """Recurrent weights [D, H, 4H] -> the tensor-core LSTM's resident image (bf16 hi for shared memory, bf16 lo for
    tensor memory); None when H is not served by that kernel (it handles H <= 128, H % 32 == 0, zero-padded to 128 units)."""
    lib = _lib.load()
    nbytes = lib.ps_lstm_packed_bytes(H, D)
    if nbytes == 0:
        return None
    out = torch.empty(nbytes, device=w_hh_t.device, dtype=torch.uint8)
    _lib.check(lib.ps_lstm_pack_weights(_req(w_hh_t, "lstm w_hh_t").data_ptr(), H, D, out.data_ptr(), _stream()), "ps_lstm_pack_weights")
    _launched()
    return out


def _set_side(d, side: str, x: Tensor, strides, pro: Prologue):
    setattr(d, side, x.data_ptr())
    setattr(d, side + "_batch_stride", strides[0])
    setattr(d, side + "_row_stride", strides[-1])
    if len(strides) == 3:
        setattr(d, side + "_mid_stride", strides[1])
    setattr(d, side + "_mode", pro.mode)
    setattr(d, side + "_act", pro.act)
    setattr(d, side + "_pa", _p(pro.a))
    setattr(d, side + "_pb", _p(pro.b))
    setattr(d, side + "_pro_batch_stride", pro.batch_stride)
    setattr(d, side + "_rowstats", _p(pro.rowstats))
    setattr(d, side + "_slope", _p(pro.slope))


def gated(a: Tensor, pro_a: Prologue, b: Optional[Tensor] = None, pro_b: Prologue = NO_PRO, *, batch: int, rows: int, C_: int,
          a_strides: Optional[tuple] = None, b_strides: Optional[tuple] = None, out: Optional[Tensor] = None,
          y_strides: Optional[tuple] = None, mid: int = 1) -> Tensor:
    """y = a' * sigmoid(b') with a' = act(pro_a(a)), b' = act(pro_b(b)) (the GatedTCN product, conv_tasnet.py:205); with
    b=None y = a' (strided transform copy).  Strides are (batch, row) - or (batch, mid, row) with a middle level of `mid`
    entries - in floats; default contiguous [batch, mid, rows, C]."""
    lib = _lib.load()
    y = out if out is not None else torch.empty(batch, mid * rows, C_, device=a.device, dtype=torch.float32)
    d = _lib.GatedDesc()
    d.batch, d.rows, d.C, d.mid = batch, rows, C_, mid
    dflt = (mid * rows * C_, rows * C_, C_)
    _set_side(d, "a", _dev(a, "gated a"), a_strides or dflt, pro_a)
    if b is not None:
        _set_side(d, "b", _dev(b, "gated b"), b_strides or dflt, pro_b)
    d.y = y.data_ptr()
    ys = y_strides or dflt
    d.y_batch_stride, d.y_row_stride = ys[0], ys[-1]
    if len(ys) == 3:
        d.y_mid_stride = ys[1]
    _lib.check(lib.ps_gated(C.byref(d), _stream()), "ps_gated")
    _launched()
    return y


def attention(qkv: Tensor, heads: int, causal: bool = False) -> Tensor:
    """qkv [B, L, 3E] (fused in-projection q | k | v) -> softmax(q k^T / sqrt(E/heads)) v, [B, L, E]."""
    lib = _lib.load()
    B, L, E3 = _req(qkv, "attention qkv").shape
    E = E3 // 3
    out = torch.empty(B, L, E, device=qkv.device, dtype=torch.float32)
    _lib.check(lib.ps_attention(qkv.data_ptr(), out.data_ptr(), B, L, E, heads, 1 if causal else 0, _stream()), "ps_attention")
    _launched()
    return out


def sdr(s1: Tensor, s2: Tensor, *, scaled: bool = True, scale_dependent: bool = False, zero_mean: bool = True,
        sdr_max: Optional[float] = None, eps: float = 1e-8) -> Tensor:
    """s1 (estimate), s2 (reference) [rows, L] -> SDR-family score in dB per row [rows] (loss/sdr.py:104-185)."""
    lib = _lib.load()
    _req(s1, "sdr s1")
    _req(s2, "sdr s2")
    if s1.shape != s2.shape or s1.dim() != 2:
        raise ValueError(f"sdr: expected two [rows, L] tensors, got {tuple(s1.shape)} and {tuple(s2.shape)}")
    rows, L = s1.shape
    out = torch.empty(rows, device=s1.device, dtype=torch.float32)
    tau = 0.0 if sdr_max is None else 10.0 ** (-sdr_max / 10.0)
    _lib.check(lib.ps_sdr(s1.data_ptr(), s2.data_ptr(), rows, L, L, L, int(scaled), int(scale_dependent), int(zero_mean), tau, eps,
                          out.data_ptr(), _stream()), "ps_sdr")
    _launched()
    return out


def film_combine(sb: Tensor, xn: Tensor) -> Tensor:
    lib = _lib.load()
    Cn = xn.shape[-1]
    y = torch.empty_like(xn)
    _lib.check(lib.ps_film_combine(_req(sb, "film sb").data_ptr(), xn.data_ptr(), y.data_ptr(), xn.numel() // Cn, Cn,
                                   _stream()), "ps_film_combine")
    _launched()
    return y


def transpose(x: Tensor) -> Tensor:
    """[B, R, C] -> [B, C, R] (the reference's [N,C,T] <-> frames-major [N,T,C])."""
    lib = _lib.load()
    x = x.contiguous()
    B, R, Cn = _req(x, "transpose").shape
    y = torch.empty(B, Cn, R, device=x.device, dtype=torch.float32)
    _lib.check(lib.ps_transpose(x.data_ptr(), y.data_ptr(), B, R, Cn, _stream()), "ps_transpose")
    _launched()
    return y
