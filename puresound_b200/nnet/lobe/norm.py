"""Normalisation layers of the separator path (mirror of puresound/nnet/lobe/norm.py).

These modules are *parameter holders with the reference's state-dict keys*
(``gamma``/``beta`` for gLN/cLN, ``weight``/``bias`` for gGN, BatchNorm1d buffers
for bN1d).  Inside the engine a norm is never a standalone pass: the producer
kernel emits its statistics and the consumer applies it on load
(``puresound_b200.nnet._fuse``).  A standalone ``forward`` is provided for cLN
(one row-norm kernel); global gLN as an isolated op is not on the hot path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops


class _LayerNorm(nn.Module):
    def __init__(self, channel_size: int):
        super().__init__()
        self.eps = 1e-8
        self.channel_size = channel_size
        self.gamma = nn.Parameter(torch.ones(channel_size))
        self.beta = nn.Parameter(torch.zeros(channel_size))


class GlobLN(_LayerNorm):
    """gLN: statistics over all of (C, T) per item (reference lobe/norm.py:20-34)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError(
            "gLN is fused into its neighbours (producer statistics + consumer prologue); "
            "call the enclosing TCN / ConvTasNet module instead"
        )


class ChanLN(_LayerNorm):
    """cLN: statistics over C for every frame (reference lobe/norm.py:37-50)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # [N, C, T]
        xt = ops.transpose(x)  # [N, T, C]
        return ops.transpose(ops.rownorm(xt, self.gamma, self.beta, self.eps))


gLN = GlobLN
cLN = ChanLN
bN1d = nn.BatchNorm1d


def gGN(channels: int) -> nn.GroupNorm:
    return nn.GroupNorm(1, channels, 1e-8)


_REGISTRY = {"gLN": gLN, "cLN": cLN, "gGN": gGN, "bN1d": bN1d}


def get_norm(name: str):
    """Same registry contract as the reference (lobe/norm.py:100-112): unknown names
    raise NameError.  iLN / bN2d belong to the out-of-scope U-Net models."""
    if name in ("iLN", "bN2d"):
        raise NotImplementedError(f"{name} is only used by the out-of-scope U-Net models (SURVEY.md section 2)")
    if name not in _REGISTRY:
        raise NameError("Could not interpret normalization identifier")
    return _REGISTRY[name]
