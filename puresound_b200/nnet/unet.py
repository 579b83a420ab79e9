"""U-Net shell with a TCN bottleneck on the B200 engine (drop-in for ``puresound.nnet.unet.Unet`` / ``UnetTcn``,
reference unet.py:13-556; SURVEY.md 8f rank 1, second half - the reference's STFT-domain TSE recipes
``tse_unet_tcn_v0 / _v0_causal / _v1``).  Same constructors, sub-module tree and ``state_dict`` keys.

Layout: activations are ``[N, T, F, C]`` (channel fastest), the reference's ``[N, C, F, T]`` re-ordered, so that a
frequency window of a frame is one contiguous run of memory.  Every 2-D convolution then is the engine's GEMM:

* ``Conv2d(kernel (kf, kt), stride (s, 1))`` after ``ZeroPad2d`` (unet.py:106-125): the input is copied into a tap buffer
  ``[N*T, F + 2*(kf//2), kt*C]`` - frequency zero-padded, the ``kt`` time taps side by side on the channel axis (slot j
  holds frame t + j - (kt - delay - 1), zero outside the utterance) - and output row ``(n, t, f')`` is ONE GEMM row of
  ``K = kf*kt*C`` contiguous floats starting at frequency ``s*f'`` (rows overlap: row stride ``s*kt*C < K``, read in place
  as the framed encoder does).
* ``ConvTranspose2d(kernel (k, tk), stride (s, 1), padding (k//2, 0), output_padding (s-k+2*(k//2), 0))``
  (unet.py:131-170) in gather form, one GEMM per output phase ``phi = f_out % s``: the taps ``kf = kappa + s*m`` that reach
  phase phi read a contiguous window of input frequencies, and the phase's rows are written interleaved
  (output row stride ``s*C_out``).  The time trim after every up layer (unet.py:529-537) only selects which frames the
  ``tk`` slots hold; the channel concatenation with the skip connection (unet.py:524) is two column blocks of the slots.

Norm + PReLU between layers: gLN statistics (global over (C, F, T) per item, lobe/norm.py:20-34) come from the GEMM
epilogue's Welford partials merged per item; BatchNorm2d (eval) folds to a per-channel affine; both are applied by the
``ps_gated`` transform-copy kernel when the activated tensor is written (it is needed twice: next layer + skip).
The bottleneck is the engine's ``TCN`` / ``GatedTCN`` stack on ``[N, T, C*F']``.
"""
from __future__ import annotations

import os

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_NONE, ACT_PRELU, PRO_AFFINE, Prologue
from ._fuse import ParamCache, prelu_slope
from .conv_tasnet import TCN, GatedTCN
from .lobe.norm import GlobLN


def _norm_cls(name: str):
    """The 2-D norms the reference's U-Net recipes use (lobe/norm.py:94-112 registry)."""
    if name not in ["gLN", "cLN", "iLN", "bN1d", "gGN", "bN2d"]:
        raise NameError("Could not interpret normalization identifier")
    if name == "gLN":
        return GlobLN
    if name == "bN2d":
        return nn.BatchNorm2d
    raise NotImplementedError(f"U-Net norm {name}: the reference's recipes use gLN and bN2d")


def _act_cls(name: str):
    if name not in ["relu", "mish", "prelu", "sigmoid", "tanh"]:
        raise NameError("Could not interpret activation identifier")
    if name != "prelu":
        raise NotImplementedError(f"U-Net activation {name}: the reference's recipes use PReLU")
    return nn.PReLU


class Unet(nn.Module):
    """reference: unet.py:13-295."""

    def __init__(
        self,
        input_type: str = "RI",
        input_dim: int = 512,
        activation_type: str = "PReLU",
        norm_type: str = "bN2d",
        dropout: float = 0.05,
        channels: Tuple = (1, 1, 8, 8, 16, 16),
        transpose_t_size: int = 2,
        skip_conv: bool = False,
        kernel_t: Tuple = (5, 1, 9, 1, 1),
        stride_t: Tuple = (1, 1, 1, 1, 1),
        dilation_t: Tuple = (1, 1, 1, 1, 1),
        kernel_f: Tuple = (1, 5, 1, 5, 1),
        stride_f: Tuple = (1, 4, 1, 4, 1),
        dilation_f: Tuple = (1, 1, 1, 1, 1),
        delay: Tuple = (0, 0, 1, 0, 0),
        multi_output: int = 1,
    ):
        super().__init__()
        assert len(kernel_t) == len(kernel_f) == len(stride_t) == len(stride_f) == len(dilation_t) == len(dilation_f)
        self.input_type, self.input_dim, self.multi_output = input_type, input_dim, multi_output
        self.activation_type, self.norm_type, self.dropout, self.skip_conv = activation_type, norm_type, dropout, skip_conv
        self.kernel_t, self.kernel_f, self.stride_t, self.stride_f = kernel_t, kernel_f, stride_t, stride_f
        self.dilation_t, self.dilation_f, self.transpose_t_size = dilation_t, dilation_f, transpose_t_size
        active_cls = _act_cls(activation_type.lower())
        norm_cls = _norm_cls(norm_type)
        self.n_cnn = len(kernel_t)
        self.channels = list(channels)
        self.kernel = list(zip(kernel_f, kernel_t))
        self.delay = delay
        self.dilation = list(zip(dilation_f, dilation_t))
        self.stride = list(zip(stride_f, stride_t))
        self.t_kernel = transpose_t_size
        if input_type.lower() == "ri":
            self.num_freq = input_dim // 2
            self.channels[0] = self.channels[0] * 2
        elif input_type.lower() == "real":
            self.num_freq = input_dim
        else:
            raise TypeError("Input feature type should be RI-concate, RI-stack or Real")
        if skip_conv or multi_output != 1:
            raise NotImplementedError("skip_conv / multi_output U-Nets are not used by the reference's recipes")
        if any(d != (1, 1) for d in self.dilation) or any(s[1] != 1 for s in self.stride):
            raise NotImplementedError("dilated / time-strided U-Net convolutions are not used by the reference's recipes")
        self.cnn_down = nn.ModuleList()
        for i in range(self.n_cnn):
            freq_pad = (self.kernel[i][0] // 2, self.kernel[i][0] // 2)
            time_pad = (self.kernel[i][1] - self.delay[i] - 1, self.delay[i])
            self.cnn_down.append(nn.Sequential(
                nn.ZeroPad2d(time_pad + freq_pad),
                nn.Conv2d(self.channels[i], self.channels[i + 1], kernel_size=self.kernel[i], stride=self.stride[i], dilation=self.dilation[i]),
                norm_cls(self.channels[i + 1]), active_cls(), nn.Dropout(self.dropout),
            ))
        self.cnn_up = nn.ModuleList()
        for i in reversed(range(self.n_cnn)):
            s, _ = self.stride[i]
            k = self.kernel[i][0]
            p = k // 2
            op = s - k + 2 * p
            layer = [nn.ConvTranspose2d(self.channels[i + 1] * 2, self.channels[i], kernel_size=(k, self.t_kernel), stride=self.stride[i],
                                        dilation=self.dilation[i], padding=(p, 0), output_padding=(op, 0))]
            if i != 0:
                layer += [norm_cls(self.channels[i]), active_cls()]
            self.cnn_up.append(nn.Sequential(*layer))
        self._cache = ParamCache()
        self.transpose_delay = False

    # ------------------------------------------------------------------ engine helpers
    def _check_eval(self):
        if self.training and self.dropout > 0:
            raise NotImplementedError("dropout > 0 in train mode is a training feature (out of scope)")

    def _activate(self, rawv, norm: Optional[nn.Module], act: Optional[nn.PReLU], N: int) -> torch.Tensor:
        """rawv = (tensor, frames per item, rows allocated per frame, valid rows per frame, C): the raw output of a layer's
        GEMMs, whose allocation carries rows that are not part of the tensor -> PReLU(norm(raw)) (or a plain copy when the
        layer has no norm) as a contiguous [N*frames, valid rows, C] tensor.  gLN statistics are taken over the valid region
        only (`ps_stats_region`), then folded to a per-item per-channel affine applied by the copy."""
        raw, Tn, R, Fv, C_ = rawv
        strides = (Tn * R * C_, R * C_, C_)
        if norm is None:
            pro = ops.NO_PRO
        elif isinstance(norm, GlobLN):
            part = ops.stats_region(raw, batch=N, mid=Tn, rows=Fv, C_=C_, strides=strides)
            scale, shift = ops.stats_finalize(part, norm.gamma, norm.beta, norm.eps, C_)
            pro = Prologue(PRO_AFFINE, ACT_PRELU, scale, shift, C_, None, prelu_slope(act))
        else:  # BatchNorm2d, eval
            if norm.training:
                raise NotImplementedError("train-mode BatchNorm couples batch items; the engine runs .eval() models only")
            scale, shift = ops.bn_fold(norm.weight, norm.bias, norm.running_mean, norm.running_var, norm.eps)
            pro = Prologue(PRO_AFFINE, ACT_PRELU, scale, shift, 0, None, prelu_slope(act))
        return ops.gated(raw, pro, batch=N, mid=Tn, rows=Fv, C_=C_, a_strides=strides).view(N * Tn, Fv, C_)

    @staticmethod
    def _stack(dst: torch.Tensor, T_dst: int, src, N: int, slot_width: int, col: int, shifts: List[int], f_lo: int):
        """dst [N*T_dst, F_pad, n_slots*slot_width] (zeros) <- src = (tensor [N*T_alloc, F, C], T_alloc, t_off, T): slot j,
        columns [col, col + C), frequency rows [f_lo, f_lo + F); destination frame t takes the source's valid frame
        t + shifts[j] (valid frames are [t_off, t_off + T) of each item's T_alloc) where that exists."""
        x, T_alloc, t_off, T = src
        Fp, W = dst.shape[1], dst.shape[2]
        F_, C_ = x.shape[1], x.shape[2]
        d4 = dst.view(N, T_dst, Fp, W)
        for j, sh in enumerate(shifts):
            t0, t1 = max(0, -sh), min(T_dst, T - sh)  # destination frames with a source frame
            c0 = j * slot_width + col
            # the buffer comes uninitialised (_tap_buffer zeroes only its padding rows): the frames of this slot that have
            # no source frame - the time padding at the edges of every item - are zeroed here, the rest is written below
            if t1 <= t0:
                d4[:, :, f_lo:f_lo + F_, c0:c0 + C_].zero_()
                continue
            if t0 > 0:
                d4[:, :t0, f_lo:f_lo + F_, c0:c0 + C_].zero_()
            if t1 < T_dst:
                d4[:, t1:, f_lo:f_lo + F_, c0:c0 + C_].zero_()
            a = x.view(-1)[(t_off + t0 + sh) * F_ * C_:]
            y = dst.view(-1)[(t0 * Fp + f_lo) * W + j * slot_width + col:]
            ops.gated(a, ops.NO_PRO, batch=N, mid=t1 - t0, rows=F_, C_=C_, a_strides=(T_alloc * F_ * C_, F_ * C_, C_),
                      out=y, y_strides=(T_dst * Fp * W, Fp * W, W))

    @staticmethod
    def _tap_buffer(N: int, T: int, Fp: int, W: int, f_lo: int, F_: int, device) -> torch.Tensor:
        """[N*T + 1, Fp, W] tap buffer (+1 frame of slack: the last window positions read past the end) with only its PADDING
        zeroed - the frequency rows outside [f_lo, f_lo + F) and the slack frame; every other element is written by _stack.
        (A zero-fill of the whole buffer per layer was 7 % of a tse_unet_tcn_v0 forward, run 53.  PS_UNET_POISON=1 fills the
        buffer with NaN first: any element neither zeroed nor written then shows up in the output - used by the tests.)"""
        buf = torch.empty(N * T + 1, Fp, W, device=device, dtype=torch.float32)
        if os.environ.get("PS_UNET_POISON") == "1":
            buf.fill_(float("nan"))
        b4 = buf[:N * T].view(N, T, Fp, W)
        if f_lo > 0:
            b4[:, :, :f_lo].zero_()
        if f_lo + F_ < Fp:
            b4[:, :, f_lo + F_:].zero_()
        buf[N * T:].zero_()
        return buf

    def _down(self, i: int, x: torch.Tensor, N: int, T: int):
        """Activated layer input x [N*T, F, C] -> raw conv output as (tensor, T, R, F', C'): one flat framed GEMM per item over
        all T*R window positions (R = padded frequency rows / stride >= F'; the trailing R - F' positions of a frame straddle
        two frames and are never read back), so tiles stay dense however few frequency rows a deep layer has."""
        conv = self.cnn_down[i][1]
        kf, kt = self.kernel[i]
        s = self.stride[i][0]
        F_, C_ = x.shape[1], x.shape[2]
        pf = kf // 2
        left = kt - self.delay[i] - 1
        # P adjacent output positions share one GEMM row (output channels (p, co), window kf + (P-1)*s rows, a position's
        # weights zero outside its own kf rows): M = P * C_out fills the 256-channel MMA of the tensor-core kernel instead
        # of zero-padding C_out = 32..128 to it, at the price of (kf + (P-1)*s) / (P*kf) of the K work.
        M1 = conv.out_channels
        F_out = (F_ + 2 * pf - kf) // s + 1
        P = 1
        while P * 2 * M1 <= 256 and P * 2 <= F_out:
            P *= 2
        step = P * s
        Fp = -(-(F_ + 2 * pf) // step) * step   # frequency rows per frame, a multiple of the row step so windows are equidistant
        W = kt * C_
        buf = self._tap_buffer(N, T, Fp, W, pf, F_, x.device)
        self._stack(buf[:N * T], T, (x, T, 0, T), N, C_, 0, [j - left for j in range(kt)], pf)
        rows_w = kf + (P - 1) * s
        R = Fp // step
        M, K = P * M1, rows_w * W

        def build():
            w = torch.zeros(P, M1, rows_w, kt, C_, device=conv.weight.device, dtype=torch.float32)
            for pp in range(P):
                w[pp, :, pp * s:pp * s + kf] = conv.weight.permute(0, 2, 3, 1)  # [C_out, C_in, kf, kt] -> [C_out, kf, kt, C_in]
            w = w.reshape(M, K).contiguous()
            bias = conv.bias.repeat(P).contiguous() if conv.bias is not None else None
            return w, bias, ops.pack_weights(w, M, K, K)

        w, bias, pk = self._cache.get(f"down{i}", [conv.weight] + ([conv.bias] if conv.bias is not None else []), build)
        y, _ = ops.gemm(buf.view(-1), w, batch=N, rows=T * R, M=M, K=K, x_batch_stride=T * Fp * W, x_row_stride=step * W, w_row_stride=K,
                        bias=bias, w_packed=pk)
        R, M = R * P, M1  # the same memory as [frames, R*P positions, C_out]
        return (y, T, R, F_out, M)

    def _up(self, i: int, xs, skip: torch.Tensor, N: int, T: int):
        """cat([x, skip]) -> raw transposed-conv output over ALL T + tk - 1 output frames, as ((tensor, To, rows allocated,
        valid rows, C_out), t_off): the reference normalises the untrimmed tensor (gLN statistics include the frames the trim
        then drops, unet.py:527-537), so the extra frames are computed and only skipped when the next layer reads.
        xs = (tensor, T_alloc, t_off, T) view of the layer input; skip [N*T, F, C].  One flat GEMM per item and output phase."""
        x = xs[0]
        conv = self.cnn_up[i][0]
        idx = self.n_cnn - 1 - i
        k, tk = self.kernel[idx][0], self.t_kernel
        s = self.stride[idx][0]
        p = k // 2
        F_, C_ = x.shape[1], x.shape[2]
        assert skip.shape[1:] == x.shape[1:], (skip.shape, x.shape)
        Cin, Cout = 2 * C_, conv.out_channels
        # phase phi = f_out % s is reached by taps kf = kappa + s*m (m < Mp) from input rows j + q - m, j = f_out // s
        phases = []
        for phi in range(s):
            kappa = (phi + p) % s
            Mp = len(range(kappa, k, s))
            phases.append((phi, kappa, Mp, (phi + p - kappa) // s))
        if any(Mp == 0 for _, _, Mp, _ in phases):
            raise NotImplementedError("transposed conv with stride > kernel")
        pad_lo = max(max(Mp - 1 - q for _, _, Mp, q in phases), 0)
        pad_hi = max(max(q for _, _, _, q in phases), 0)
        W = tk * Cin
        Rb = F_ + pad_lo + pad_hi
        To = T + tk - 1  # output frame t_o = t_i + kt: slot kt holds input frame t_o - kt
        buf = self._tap_buffer(N, To, Rb, W, pad_lo, F_, x.device)
        shifts = [-j for j in range(tk)]
        self._stack(buf[:N * To], To, xs, N, Cin, 0, shifts, pad_lo)
        self._stack(buf[:N * To], To, (skip, T, 0, T), N, Cin, C_, shifts, pad_lo)
        # All s phases in ONE GEMM: output channels (phi, co) over the union of the phases' windows (rows j + lo .. j + hi of
        # the tap buffer; a phase's weights are zero on rows it does not reach).  M = s * C_out fills the 256-channel MMA
        # better than s launches of C_out, and the row (j, phi, co) IS the interleaved output layout f_out = s*j + phi.
        lo = min(q - (Mp - 1) for _, _, Mp, q in phases)
        hi = max(q for _, _, _, q in phases)
        rows_w = hi - lo + 1
        K, M = rows_w * W, s * Cout

        def build():
            w = torch.zeros(s, Cout, rows_w, tk, Cin, device=conv.weight.device, dtype=torch.float32)
            for phi, kappa, Mp, q in phases:
                for o in range(rows_w):
                    m = q - (lo + o)
                    if 0 <= m < Mp:
                        w[phi, :, o] = conv.weight[:, :, kappa + s * m, :].permute(1, 2, 0)  # [Cin, Cout, tk] -> [Cout, tk, Cin]
            bias = conv.bias.repeat(s).contiguous() if conv.bias is not None else None
            w = w.reshape(M, K).contiguous()
            return w, bias, ops.pack_weights(w, M, K, K)

        w, bias, pk = self._cache.get(f"up{i}", [conv.weight] + ([conv.bias] if conv.bias is not None else []), build)
        y, _ = ops.gemm(buf.view(-1)[(pad_lo + lo) * W:], w, batch=N, rows=To * Rb, M=M, K=K, x_batch_stride=To * Rb * W, x_row_stride=W,
                        w_row_stride=K, bias=bias, w_packed=pk)
        y = y.view(N * To, s * Rb, Cout)
        return (y, To, s * Rb, s * F_, Cout), ((tk - 1) if self.transpose_delay else 0)

    # ------------------------------------------------------------------ forward
    def _to_cl4(self, x: torch.Tensor) -> torch.Tensor:
        """frames-major [N, T, C_in] -> [N*T, F, ch]: re | im halves become two channels (unet.py:226-228)."""
        N, T, Cx = x.shape
        if self.input_type.lower() == "ri":
            return x.view(N, T, 2, Cx // 2).permute(0, 1, 3, 2).reshape(N * T, Cx // 2, 2).contiguous()
        return x.reshape(N * T, Cx, 1)

    def _from_cl4(self, ys, N: int, T: int) -> torch.Tensor:
        y, T_alloc, t_off, _ = ys
        y = y.view(N, T_alloc, y.shape[1], y.shape[2])[:, t_off:t_off + T]
        if self.input_type.lower() == "ri":
            return y.permute(0, 1, 3, 2).reshape(N, T, -1).contiguous()
        return y.reshape(N, T, -1).contiguous()

    def _bottleneck(self, x: torch.Tensor, N: int, T: int, dvec) -> torch.Tensor:
        return x

    def forward_cl(self, x: torch.Tensor, dvec: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [N, T, C] frames-major -> [N, T, C]."""
        self._check_eval()
        N, T, _ = x.shape
        cur = self._to_cl4(x.contiguous())
        skips = []
        for i in range(self.n_cnn):
            cur = self._activate(self._down(i, cur, N, T), self.cnn_down[i][2], self.cnn_down[i][3], N)
            skips.append(cur)
        cur = (self._bottleneck(cur, N, T, dvec), T, 0, T)
        for i in range(self.n_cnn):
            rawv, t_off = self._up(i, cur, skips[-i - 1], N, T)
            last = len(self.cnn_up[i]) == 1  # the output layer is linear (unet.py:154-170)
            act = self._activate(rawv, None if last else self.cnn_up[i][1], None if last else self.cnn_up[i][2], N)
            cur = (act, rawv[1], t_off, T)
        return self._from_cl4(cur, N, T)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [N, C, T] -> [N, C, T]  (reference unet.py:219-273)."""
        return ops.transpose(self.forward_cl(ops.transpose(x)))

    @property
    def get_args(self) -> Dict:
        return {
            "input_type": self.input_type, "input_dim": self.input_dim, "activation_type": self.activation_type,
            "norm_type": self.norm_type, "dropout": self.dropout, "channels": self.channels,
            "transpose_t_size": self.transpose_t_size, "skip_conv": self.skip_conv, "kernel_t": self.kernel_t,
            "stride_t": self.stride_t, "dilation_t": self.dilation_t, "kernel_f": self.kernel_f, "stride_f": self.stride_f,
            "dilation_f": self.dilation_f, "delay": self.delay, "multi_output": self.multi_output,
        }


class UnetTcn(Unet):
    """reference: unet.py:298-556."""

    def __init__(
        self,
        embed_dim: int = 0,
        embed_norm: bool = False,
        input_type: str = "RI",
        input_dim: int = 512,
        activation_type: str = "PReLU",
        norm_type: str = "bN2d",
        dropout: float = 0.05,
        channels: Tuple = (1, 1, 8, 8, 16, 16),
        transpose_t_size: int = 2,
        transpose_delay: bool = False,
        skip_conv: bool = False,
        kernel_t: Tuple = (5, 1, 9, 1, 1),
        stride_t: Tuple = (1, 1, 1, 1, 1),
        dilation_t: Tuple = (1, 1, 1, 1, 1),
        kernel_f: Tuple = (1, 5, 1, 5, 1),
        stride_f: Tuple = (1, 4, 1, 4, 1),
        dilation_f: Tuple = (1, 1, 1, 1, 1),
        delay: Tuple = (0, 0, 1, 0, 0),
        tcn_layer: str = "normal",
        tcn_kernel: int = 3,
        tcn_dim: int = 256,
        tcn_dilated_basic: int = 2,
        per_tcn_stack: int = 5,
        repeat_tcn: int = 4,
        tcn_with_embed: List = [1, 0, 0, 0, 0],
        tcn_use_film: bool = False,
        tcn_norm: str = "gLN",
        dconv_norm: str = "gGN",
        causal: bool = False,
    ):
        super().__init__(input_type, input_dim, activation_type, norm_type, dropout, channels, transpose_t_size, skip_conv,
                         kernel_t, stride_t, dilation_t, kernel_f, stride_f, dilation_f, delay)
        self.embed_dim, self.embed_norm = embed_dim, embed_norm
        self.tcn_layer, self.tcn_dim, self.tcn_kernel = tcn_layer, tcn_dim, tcn_kernel
        self.per_tcn_stack, self.repeat_tcn, self.tcn_dilated_basic = per_tcn_stack, repeat_tcn, tcn_dilated_basic
        self.tcn_with_embed, self.tcn_norm, self.dconv_norm = tcn_with_embed, tcn_norm, dconv_norm
        self.tcn_use_film, self.causal, self.transpose_delay = tcn_use_film, causal, transpose_delay
        temporal_input_dim = self.num_freq
        for stride, _ in self.stride:
            temporal_input_dim = temporal_input_dim // stride + (1 if temporal_input_dim % stride else 0)
        temporal_input_dim *= self.channels[-1]
        if self.tcn_layer.lower() not in ("normal", "gated"):
            raise NameError
        gated = self.tcn_layer.lower() == "gated"
        assert per_tcn_stack == len(tcn_with_embed)
        self.tcn_list = nn.ModuleList()
        for _ in range(repeat_tcn):
            stack = []
            for i in range(per_tcn_stack):
                emb = embed_dim if tcn_with_embed[i] else 0
                kw = dict(kernel=tcn_kernel, dilation=tcn_dilated_basic ** i, emb_dim=emb, causal=causal, tcn_norm=tcn_norm)
                if gated:
                    stack.append(GatedTCN(temporal_input_dim, tcn_dim, use_film=tcn_use_film if tcn_with_embed[i] else False, **kw))
                else:
                    stack.append(TCN(temporal_input_dim, tcn_dim, dconv_norm=dconv_norm, **kw))
            self.tcn_list.append(nn.ModuleList(stack))

    def _bottleneck(self, x: torch.Tensor, N: int, T: int, dvec) -> torch.Tensor:
        """[N*T, F', C] -> TCN stack on [N, T, C*F'] (channel-major flattening, unet.py:489-499) -> back."""
        if self.embed_norm and dvec is not None:
            dvec = ops.l2normalize(dvec.contiguous())
        F_, C_ = x.shape[1], x.shape[2]
        y = x.view(N, T, F_, C_).permute(0, 1, 3, 2).reshape(N, T, C_ * F_).contiguous()
        for stack in self.tcn_list:
            for i, blk in enumerate(stack):
                y = blk.forward_cl(y, dvec if self.tcn_with_embed[i] else None)
        return y.view(N, T, C_, F_).permute(0, 1, 3, 2).reshape(N * T, F_, C_).contiguous()

    @torch.no_grad()
    def forward(self, x: torch.Tensor, dvec: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [N, C, T], dvec [N, E] -> [N, C, T]  (reference unet.py:454-517)."""
        return ops.transpose(self.forward_cl(ops.transpose(x), dvec))

    @property
    def get_args(self) -> Dict:
        a = super().get_args
        a.pop("multi_output")
        a.update({
            "transpose_delay": self.transpose_delay, "embed_dim": self.embed_dim, "embed_norm": self.embed_norm,
            "tcn_norm": self.tcn_norm, "dconv_norm": self.dconv_norm, "tcn_layer": self.tcn_layer, "tcn_dim": self.tcn_dim,
            "tcn_kernel": self.tcn_kernel, "tcn_dilated_basic": self.tcn_dilated_basic, "repeat_tcn": self.repeat_tcn,
            "per_tcn_stack": self.per_tcn_stack, "tcn_with_embed": self.tcn_with_embed, "tcn_use_film": self.tcn_use_film,
            "causal": self.causal,
        })
        return a
