"""Conv-TasNet TCN stack on the B200 engine.

Drop-in for ``puresound.nnet.conv_tasnet.TCN`` / ``GatedTCN`` / ``ConvTasNet`` (tcn_layer "normal" or "gated"):
same constructor signatures, same sub-module tree and state-dict keys, same
``forward(x[N,C,T], dvec[N,E]) -> [N,C,T]``.  One TCN block is four fused kernels
(+ three tiny statistics merges for gLN/gGN):

    GEMM(W_in)  [+ per-item embedding bias]            -> u1, Welford partials
    depthwise   [prologue PReLU(norm1(u1))] + bias      -> u2, partials
    GEMM(W_pw)  [prologue PReLU(norm2(u2))] + bias      -> u3, partials
    GEMM(W_out) [prologue PReLU(norm3(u3))] + bias + x  -> y

so each [N,T,C] tensor is written once and read once (the residual twice).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import ops
from ._fuse import ParamCache, norm_prologue, prelu_slope, stats_request
from .lobe.cnn import DepthwiseSeparableConv1d
from .lobe.norm import get_norm


class TCN(nn.Module):
    """reference: conv_tasnet.py:11-90."""

    def __init__(
        self,
        in_channels: int,
        hid_channels: int,
        kernel: int,
        dilation: int,
        dropout: float = 0.0,
        emb_dim: int = 0,
        causal: bool = False,
        tcn_norm: str = "gLN",
        dconv_norm: str = "gGN",
    ) -> None:
        super().__init__()
        norm = get_norm(tcn_norm)
        self.in_conv = nn.Sequential(
            nn.Conv1d(in_channels + emb_dim, hid_channels, kernel_size=1, bias=False, groups=1), norm(hid_channels), nn.PReLU()
        )
        self.dconv = nn.Sequential(
            DepthwiseSeparableConv1d(hid_channels, hid_channels, None, kernel=kernel, dilation=dilation, skip=False, causal=causal, norm_cls=dconv_norm),
            nn.Dropout(p=dropout),
        )
        self.out_conv = nn.Conv1d(hid_channels, in_channels, kernel_size=1, stride=1)
        self.in_channels, self.hid_channels, self.emb_dim = in_channels, hid_channels, emb_dim
        self.kernel, self.dilation, self.causal = kernel, dilation, causal
        self._cache = ParamCache()

    def _packed(self, tag: str, w: torch.Tensor, M: int, K: int, ld: int):
        """bf16 hi/lo split + swizzled tile image of a 1x1-conv weight for the tcgen05 GEMM (None if not eligible)."""
        return self._cache.get(tag, [w], lambda: ops.pack_weights(w, M, K, ld))

    # ---- engine path: frames-major ----
    def forward_cl(self, x: torch.Tensor, embed: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [N, T, C] frames-major, embed [N, E] (already L2-normalised if requested)."""
        if self.training and self.dconv[1].p > 0:
            raise NotImplementedError("dropout > 0 in train mode is a training feature (out of scope)")
        C, H, E = self.in_channels, self.hid_channels, self.emb_dim
        dsc = self.dconv[0]
        w_in = self.in_conv[0].weight.view(H, C + E)
        bias_item = None
        if embed is not None:
            if E == 0:
                raise ValueError("this TCN block was built with emb_dim=0")
            # cat(x, repeat(embed)) through W_in  ==  W_in[:, :C] x + (W_in[:, C:] embed) as a per-item bias
            eb, _ = ops.gemm(embed, w_in[:, C:], batch=1, rows=embed.shape[0], M=H, K=E, x_batch_stride=0,
                             x_row_stride=E, w_row_stride=C + E)
            bias_item = eb.view(embed.shape[0], H)
        elif E != 0:
            raise ValueError("this TCN block expects a conditioning embedding")
        n1, n2, n3 = self.in_conv[1], dsc.depthwise[1], dsc.pointwise[1]
        u1, p1 = ops.linear(x, w_in, K=C, w_row_stride=C + E, bias_batch=bias_item, **stats_request(n1),
                            w_packed=self._packed("in", self.in_conv[0].weight, H, C, C + E))
        pro1 = norm_prologue(n1, u1, p1, prelu_slope(self.in_conv[2]))
        dw = dsc.depthwise[0]
        u2, p2 = ops.dwconv(u1, dw.weight.view(H, self.kernel), dw.bias, self.kernel, self.dilation, self.causal, pro1,
                            **stats_request(n2))
        pro2 = norm_prologue(n2, u2, p2, prelu_slope(dsc.depthwise[2]))
        pw = dsc.pointwise[0]
        u3, p3 = ops.linear(u2, pw.weight.view(H, H), pro=pro2, bias=pw.bias, **stats_request(n3),
                            w_packed=self._packed("pw", pw.weight, H, H, H))
        pro3 = norm_prologue(n3, u3, p3, prelu_slope(dsc.pointwise[2]))
        y, _ = ops.linear(u3, self.out_conv.weight.view(C, H), pro=pro3, bias=self.out_conv.bias, residual=x,
                          w_packed=self._packed("out", self.out_conv.weight, C, H, H))
        return y

    # ---- reference-layout API ----
    @torch.no_grad()
    def forward(self, x: torch.Tensor, embed: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [N, C, T], embed [N, E] -> [N, C, T]  (reference conv_tasnet.py:67-90)."""
        return ops.transpose(self.forward_cl(ops.transpose(x), embed))


class GatedTCN(nn.Module):
    """Gated TCN block (reference conv_tasnet.py:93-215): 1x1 in_conv, two dense dilated k-tap convs
    ("left": norm + PReLU, "right": norm + PReLU + sigmoid, optionally conditioned on a speaker embedding by concatenation
    or FiLM), their product, 1x1 out_conv + residual.

    On the engine: in_conv writes straight into a zero-padded frames-major buffer [N, T + 2*padd, H (+E)]; a k-tap dense
    dilated conv is k GEMM launches over row-shifted views of that buffer, chained through the residual input (tap j reads
    rows t + j*dilation); the concatenated embedding is a block of extra columns filled only in the un-padded rows, which
    reproduces the reference's zero padding of the concatenated tensor at the edges exactly; both branch norms + PReLUs and
    the sigmoid ride in the `ps_gated` product kernel, so normalised tensors are never written; out_conv reads the first T
    rows (the causal trim, conv_tasnet.py:209-210, commutes with the 1x1 conv)."""

    def __init__(
        self,
        in_channels: int,
        hid_channels: int,
        kernel: int,
        dilation: int,
        dropout: float = 0.0,
        emb_dim: int = 0,
        causal: bool = False,
        tcn_norm: str = "gLN",
        use_film: bool = False,
    ):
        super().__init__()
        self.causal = causal
        self.padd = (kernel - 1) * dilation // 2 if not causal else (kernel - 1) * dilation
        self.tcn_norm = tcn_norm
        norm_cls = get_norm(tcn_norm)
        self.use_film = use_film
        self.in_conv = nn.Conv1d(in_channels, hid_channels, kernel_size=1, bias=False, groups=1)
        self.left_conv = nn.Sequential(
            nn.Conv1d(hid_channels, hid_channels, kernel_size=kernel, dilation=dilation, bias=False, padding=self.padd, groups=1),
            norm_cls(hid_channels), nn.PReLU(), nn.Dropout(p=dropout),
        )
        if not self.use_film:
            right_in_dim = hid_channels + emb_dim
        else:
            self.cond_scale = nn.Conv1d(emb_dim, hid_channels, kernel_size=1, bias=False)
            self.cond_bias = nn.Conv1d(emb_dim, hid_channels, kernel_size=1, bias=False)
            right_in_dim = hid_channels
        self.right_conv = nn.Sequential(
            nn.Conv1d(right_in_dim, hid_channels, kernel_size=kernel, dilation=dilation, bias=False, padding=self.padd, groups=1),
            norm_cls(hid_channels), nn.PReLU(), nn.Dropout(p=dropout), nn.Sigmoid(),
        )
        self.out_conv = nn.Conv1d(hid_channels, in_channels, kernel_size=1, bias=False, groups=1)
        self.in_channels, self.hid_channels, self.emb_dim = in_channels, hid_channels, emb_dim
        self.kernel, self.dilation = kernel, dilation
        self._cache = ParamCache()

    def _taps(self, tag: str, conv: nn.Conv1d):
        """Per-tap [H, K_in] matrices of a k-tap conv weight [H, K_in, k] and their tcgen05 images."""
        def build():
            w = [conv.weight[:, :, j].contiguous() for j in range(self.kernel)]
            return w, [ops.pack_weights(wj, wj.shape[0], wj.shape[1], wj.shape[1]) for wj in w]
        return self._cache.get(tag, [conv.weight], build)

    def _dilated(self, buf: torch.Tensor, rows_in: int, width: int, K: int, taps, norm: nn.Module, N: int, L_out: int):
        """sum_j W_j . buf[n, t + j*dilation, :K] for t < L_out -> (raw [N, L_out, H], statistics for `norm`)."""
        w, packed = taps
        flat = buf.view(-1)
        acc, part = None, None
        for j in range(self.kernel):
            last = j == self.kernel - 1
            acc, part = ops.gemm(flat[j * self.dilation * width:], w[j], batch=N, rows=L_out, M=self.hid_channels, K=K,
                                 x_batch_stride=rows_in * width, x_row_stride=width, w_row_stride=K, residual=acc,
                                 w_packed=packed[j], **(stats_request(norm) if last else {}))
        return acc, part

    def forward_cl(self, x: torch.Tensor, embed: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [N, T, C] frames-major, embed [N, E] (already L2-normalised if requested)."""
        if self.training and self.left_conv[3].p > 0:
            raise NotImplementedError("dropout > 0 in train mode is a training feature (out of scope)")
        N, T, C = x.shape
        H, E, p = self.hid_channels, self.emb_dim, self.padd
        if embed is None and E != 0 and not self.use_film:
            raise ValueError("this GatedTCN block expects a conditioning embedding")
        if embed is not None and E == 0:
            raise ValueError("this GatedTCN block was built with emb_dim=0")
        concat = embed is not None and not self.use_film
        width = H + (E if concat else 0)
        rows_in = T + 2 * p
        L_out = rows_in - (self.kernel - 1) * self.dilation
        if L_out < T or (not self.causal and L_out != T):
            # odd (kernel-1)*dilation without causal trim: the reference fails on `x + res` (conv_tasnet.py:213) - fail as loudly
            raise RuntimeError(f"The size of tensor a ({L_out}) must match the size of tensor b ({T}) at non-singleton dimension 2")
        # only the pad rows are zeroed: the un-padded rows are written in full below (in_conv, the embedding columns) - a
        # zero-fill of the whole [N, T + 2p, width] buffer per block was 9 % of the gated TSE step
        xp = torch.empty(N, rows_in, width, device=x.device, dtype=torch.float32)
        xp[:, :p].zero_()
        xp[:, p + T:].zero_()
        inner = xp.view(-1)[p * width:]  # first un-padded row
        ops.linear(x, self.in_conv.weight.view(H, C), out=inner, y_strides=(rows_in * width, width),
                   w_packed=self._cache.get("in", [self.in_conv.weight], lambda: ops.pack_weights(self.in_conv.weight.view(H, C), H, C, C)))
        right_in = xp
        if concat:
            xp[:, p:p + T, H:] = embed.unsqueeze(1)  # cat(x, repeat(embed)) (conv_tasnet.py:191-194); pad rows stay zero
        elif embed is not None:
            # FiLM (conv_tasnet.py:196-200): x_r = scale_n * x + bias_n, written into its own zero-padded buffer
            wsb = self._cache.get("film", [self.cond_scale.weight, self.cond_bias.weight],
                                  lambda: torch.cat([self.cond_scale.weight.view(H, E), self.cond_bias.weight.view(H, E)], 0).contiguous())
            sb, _ = ops.gemm(embed.contiguous(), wsb, batch=1, rows=N, M=2 * H, K=E, x_batch_stride=0, x_row_stride=E, w_row_stride=E)
            sb = sb.view(N, 2 * H)
            right_in = torch.empty(N, rows_in, H, device=x.device, dtype=torch.float32)
            right_in[:, :p].zero_()
            right_in[:, p + T:].zero_()
            film = ops.Prologue(ops.PRO_AFFINE, ops.ACT_NONE, sb[:, :H], sb[:, H:], 2 * H)
            ops.gated(inner, film, batch=N, rows=T, C_=H, a_strides=(rows_in * width, width),
                      out=right_in.view(-1)[p * H:], y_strides=(rows_in * H, H))
        nl, nr = self.left_conv[1], self.right_conv[1]
        left, pl = self._dilated(xp, rows_in, width, H, self._taps("left", self.left_conv[0]), nl, N, L_out)
        right, pr = self._dilated(right_in, rows_in, right_in.shape[-1], right_in.shape[-1], self._taps("right", self.right_conv[0]), nr, N, L_out)
        z = ops.gated(left, norm_prologue(nl, left, pl, prelu_slope(self.left_conv[2])),
                      right, norm_prologue(nr, right, pr, prelu_slope(self.right_conv[2])), batch=N, rows=L_out, C_=H)
        y, _ = ops.gemm(z, self.out_conv.weight.view(C, H), batch=N, rows=T, M=C, K=H, x_batch_stride=L_out * H, x_row_stride=H,
                        w_row_stride=H, residual=x,
                        w_packed=self._cache.get("out", [self.out_conv.weight], lambda: ops.pack_weights(self.out_conv.weight.view(C, H), C, H, H)))
        return y

    @torch.no_grad()
    def forward(self, x: torch.Tensor, embed: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [N, C, T], embed [N, E] -> [N, C, T]  (reference conv_tasnet.py:174-215)."""
        return ops.transpose(self.forward_cl(ops.transpose(x), embed))


class ConvTasNet(nn.Module):
    """reference: conv_tasnet.py:218-377 (encoder/decoder live in the task wrapper)."""

    def __init__(
        self,
        input_dim: int = 512,
        embed_dim: int = 256,
        embed_norm: bool = False,
        tcn_layer: str = "normal",
        tcn_kernel: int = 3,
        tcn_dim: int = 256,
        tcn_dilated_basic: int = 2,
        per_tcn_stack: int = 5,
        repeat_tcn: int = 4,
        tcn_with_embed: List = [1, 0, 0, 0, 0],
        tcn_norm: str = "gLN",
        dconv_norm: str = "gGN",
        causal: bool = False,
    ):
        super().__init__()
        self.input_dim, self.embed_dim, self.embed_norm = input_dim, embed_dim, embed_norm
        self.tcn_layer, self.tcn_dim, self.tcn_kernel = tcn_layer, tcn_dim, tcn_kernel
        self.per_tcn_stack, self.repeat_tcn, self.tcn_dilated_basic = per_tcn_stack, repeat_tcn, tcn_dilated_basic
        self.tcn_with_embed, self.tcn_norm, self.dconv_norm, self.causal = tcn_with_embed, tcn_norm, dconv_norm, causal
        if tcn_layer.lower() not in ("normal", "gated"):
            raise NameError
        gated = tcn_layer.lower() == "gated"  # GatedTCN ignores dconv_norm (conv_tasnet.py:296-331)
        assert per_tcn_stack == len(tcn_with_embed)
        self.tcn_list = nn.ModuleList()
        for _ in range(repeat_tcn):
            stack = []
            for i in range(per_tcn_stack):
                kw = dict(kernel=tcn_kernel, dilation=tcn_dilated_basic ** i, emb_dim=embed_dim if tcn_with_embed[i] else 0,
                          causal=causal, tcn_norm=tcn_norm)
                stack.append(GatedTCN(input_dim, tcn_dim, **kw) if gated else TCN(input_dim, tcn_dim, dconv_norm=dconv_norm, **kw))
            self.tcn_list.append(nn.ModuleList(stack))

    def forward_cl(self, x: torch.Tensor, dvec: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.embed_norm and dvec is not None:
            dvec = ops.l2normalize(dvec.contiguous())
        for stack in self.tcn_list:
            for i, blk in enumerate(stack):
                x = blk.forward_cl(x, dvec if self.tcn_with_embed[i] else None)
        return x

    @torch.no_grad()
    def forward(self, x: torch.Tensor, dvec: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [N, C, T], dvec [N, embed_dim] -> mask logits [N, C, T]."""
        return ops.transpose(self.forward_cl(ops.transpose(x), dvec))

    @property
    def get_args(self) -> Dict:
        return {
            "input_dim": self.input_dim, "embed_dim": self.embed_dim, "embed_norm": self.embed_norm,
            "tcn_norm": self.tcn_norm, "dconv_norm": self.dconv_norm, "tcn_layer": self.tcn_layer,
            "tcn_dim": self.tcn_dim, "tcn_kernel": self.tcn_kernel, "tcn_dilated_basic": self.tcn_dilated_basic,
            "repeat_tcn": self.repeat_tcn, "per_tcn_stack": self.per_tcn_stack,
            "tcn_with_embed": self.tcn_with_embed, "causal": self.causal,
        }
