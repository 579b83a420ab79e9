"""Conv-TasNet TCN stack on the B200 engine.

Drop-in for ``puresound.nnet.conv_tasnet.TCN`` / ``ConvTasNet`` (tcn_layer="normal"):
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
        if tcn_layer.lower() == "gated":
            raise NotImplementedError("GatedTCN is a 'next' row of the scope table (SURVEY.md 8f)")
        if tcn_layer.lower() != "normal":
            raise NameError
        assert per_tcn_stack == len(tcn_with_embed)
        self.tcn_list = nn.ModuleList()
        for _ in range(repeat_tcn):
            stack = [
                TCN(input_dim, tcn_dim, kernel=tcn_kernel, dilation=tcn_dilated_basic ** i,
                    emb_dim=embed_dim if tcn_with_embed[i] else 0, causal=causal, tcn_norm=tcn_norm, dconv_norm=dconv_norm)
                for i in range(per_tcn_stack)
            ]
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
