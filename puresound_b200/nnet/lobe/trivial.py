"""Conditioning and segmentation lobes on the B200 engine (drop-ins for
``Magnitude``, ``FiLM`` and ``SplitMerge`` of puresound/nnet/lobe/trivial.py)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn

from ... import ops
from .._fuse import ParamCache


class Magnitude(nn.Module):
    """sqrt(re^2 + im^2 + 1e-8) on channel halves (reference lobe/trivial.py:21-58)."""

    def __init__(self, drop_first: bool = True, log1p: bool = False) -> None:
        super().__init__()
        self.drop_first = drop_first
        self.log1p = log1p

    def forward_cl(self, x: torch.Tensor) -> torch.Tensor:
        """[N, T, 2F] -> [N, T, F - drop_first]"""
        return ops.magnitude(x, self.drop_first, self.log1p)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 4:  # [N, F, T, 2] -> channel-cat [N, 2F, T]
            x = torch.cat([x[..., 0], x[..., 1]], dim=1)
        elif x.dim() != 3:
            raise TypeError
        return ops.transpose(self.forward_cl(ops.transpose(x)))


class FiLM(nn.Module):
    """Feature-wise linear modulation (reference lobe/trivial.py:129-167):
    y = (W_s [x~; e]) * x~ + (W_b [x~; e]),  x~ = LayerNorm_C(x).
    One row-norm kernel, one stacked GEMM (M = 2C, the embedding columns folded into a
    per-item bias), one combine kernel."""

    def __init__(self, feats_size: int, embed_size: int, input_norm: bool = True):
        super().__init__()
        self.cond_scale = nn.Conv1d(feats_size + embed_size, feats_size, kernel_size=1, bias=False)
        self.cond_bias = nn.Conv1d(feats_size + embed_size, feats_size, kernel_size=1, bias=False)
        self.inp_norm = input_norm
        if input_norm:
            self.norm = nn.LayerNorm(feats_size)
        self._cache = ParamCache()

    def forward_cl(self, x: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
        """x [N, R, C] frames-major (R rows share item n's embedding), cond [N, E]."""
        N, R, Cn = x.shape
        E = cond.shape[1]
        xn = ops.rownorm(x, self.norm.weight, self.norm.bias, self.norm.eps) if self.inp_norm else x
        w = self._cache.get("sb", [self.cond_scale.weight, self.cond_bias.weight],
                            lambda: torch.cat([self.cond_scale.weight.view(Cn, Cn + E), self.cond_bias.weight.view(Cn, Cn + E)], 0).contiguous())
        eb, _ = ops.gemm(cond.contiguous(), w[:, Cn:], batch=1, rows=N, M=2 * Cn, K=E, x_batch_stride=0, x_row_stride=E, w_row_stride=Cn + E)
        sb, _ = ops.linear(xn, w, K=Cn, w_row_stride=Cn + E, bias_batch=eb.view(N, 2 * Cn))
        return ops.film_combine(sb, xn)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, condition: torch.Tensor) -> torch.Tensor:
        """x [N, C, T], condition [N, E] -> [N, C, T]"""
        return ops.transpose(self.forward_cl(ops.transpose(x), condition))


def overlap_geometry(T: int, K: int) -> Tuple[int, int]:
    """(rest, S) of the 50%-overlap segmentation (reference lobe/trivial.py:186-191)."""
    s = K // 2
    rest = K - (s + T % K) % K
    return rest, 2 * ((T + rest + s) // K)


class SplitMerge(nn.Module):
    """2S process: segmentation and stitching (reference lobe/trivial.py:170-241)."""

    def __init__(self, seg_size: int, seg_overlap: bool = True):
        super().__init__()
        self.seg_size = seg_size
        self.seg_overlap = seg_overlap

    @staticmethod
    @torch.no_grad()
    def split(x: torch.Tensor, seg_size: int):
        """[N, C, T] -> ([N, S, K, C], rest)"""
        T = x.shape[2]
        rest, S = overlap_geometry(T, seg_size)
        return ops.segment(ops.transpose(x), seg_size, S, True), rest

    @staticmethod
    @torch.no_grad()
    def merge(x: torch.Tensor, rest: int):
        """[N, S, K, C] -> [N, C, T]"""
        N, S, K, Cn = x.shape
        T = (S // 2) * K - K // 2 - rest
        return ops.transpose(ops.merge(x.contiguous(), T, True))
