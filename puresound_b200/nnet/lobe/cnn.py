"""Depthwise-separable Conv1d holder (mirror of puresound/nnet/lobe/cnn.py:9-106).

Keeps the reference's sub-module tree (``depthwise`` / ``pointwise`` Sequentials of
Conv1d, norm, PReLU) so state-dict keys match; the arithmetic runs in
``TCN.forward_cl`` as fused kernels.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .norm import get_norm


class DepthwiseSeparableConv1d(nn.Module):
    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        hid_channels: Optional[int] = None,
        norm_cls: str = "gGN",
        kernel: int = 3,
        stride: int = 1,
        dilation: int = 1,
        skip: bool = False,
        causal: bool = False,
    ) -> None:
        super().__init__()
        if hid_channels is not None or skip or stride != 1:
            raise NotImplementedError("the separator path uses hid_channels=None, skip=False, stride=1 (conv_tasnet.py:52-61)")
        self.skip = skip
        self.transform = False
        self.causal = causal
        if causal:
            # same guard as the reference (lobe/cnn.py:40-44): global norms would leak the future
            assert norm_cls not in ["gLN", "gGN"], "Conflict setting between normalization layer and causal operation."
        norm = get_norm(norm_cls)
        self.hid_channels = in_channels
        self.kernel = kernel
        self.dilation = dilation
        self.padding = (kernel - 1) * dilation if causal else ((kernel - 1) // 2) * dilation
        self.depthwise = nn.Sequential(
            nn.Conv1d(in_channels, in_channels, kernel_size=kernel, stride=1, dilation=dilation, padding=self.padding, groups=in_channels),
            norm(in_channels),
            nn.PReLU(),
        )
        self.pointwise = nn.Sequential(nn.Conv1d(in_channels, out_channels, kernel_size=1, stride=1), norm(out_channels), nn.PReLU())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("run through TCN.forward (its kernels span the block boundary)")
