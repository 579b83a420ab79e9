"""Attentive statistics pooling on the B200 engine (drop-in for
``puresound.nnet.lobe.pooling.AttentiveStatisticsPooling``, lobe/pooling.py:58-126).

Two GEMMs (ReLU in the first epilogue; eval-BatchNorm + tanh in the second
prologue) and one pooling kernel that does the softmax over frames and the
weighted mean / standard deviation.  ``lengths`` is not supported: every caller in
the reference passes ``None`` (all frames valid).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ... import ops
from ...ops import ACT_RELU, ACT_TANH, PRO_AFFINE, Prologue
from .._fuse import ParamCache


class AttentiveStatisticsPooling(nn.Module):
    def __init__(self, channels: int, attention_channels: int = 128):
        super().__init__()
        self.eps = 1e-12
        self.tdnn = nn.Sequential(nn.Conv1d(channels, attention_channels, kernel_size=1, dilation=1), nn.ReLU(), nn.BatchNorm1d(attention_channels))
        self.tanh = nn.Tanh()
        self.conv = nn.Conv1d(attention_channels, channels, kernel_size=1)
        self._cache = ParamCache()

    def forward_cl(self, x: torch.Tensor) -> torch.Tensor:
        """x [N, T, C] -> [N, 2C] = cat(mean, std)."""
        bn = self.tdnn[2]
        if bn.training:
            raise NotImplementedError("train-mode BatchNorm couples batch items; call .eval() first (egs/ns/main.py:113)")
        c1, c2 = self.tdnn[0], self.conv
        A, Cn = c1.out_channels, c1.in_channels
        # both 1x1 convs on tcgen05 when the channel counts allow (multiples of 32 / 64), eval-BatchNorm + tanh applied on load
        pk1 = self._cache.get("tdnn", [c1.weight], lambda: ops.pack_weights(c1.weight.view(A, Cn), A, Cn, Cn))
        pk2 = self._cache.get("conv", [c2.weight], lambda: ops.pack_weights(c2.weight.view(Cn, A), Cn, A, A))
        a1, _ = ops.linear(x, c1.weight.view(A, Cn), bias=c1.bias, epi_act=ACT_RELU, w_packed=pk1)
        scale, shift = ops.bn_fold(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
        logits, _ = ops.linear(a1, c2.weight.view(Cn, A), pro=Prologue(PRO_AFFINE, ACT_TANH, scale, shift, 0), bias=c2.bias, w_packed=pk2)
        return ops.asp_pool(x, logits)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, lengths: Optional[torch.Tensor] = None, return_weight: bool = False):
        """x [N, C, L] -> [N, 2C, 1]"""
        if lengths is not None or return_weight:
            raise NotImplementedError("lengths / return_weight are never used on the separator path")
        return self.forward_cl(ops.transpose(x)).unsqueeze(2)
