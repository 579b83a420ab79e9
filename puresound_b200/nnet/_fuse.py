"""Host-side fusion helpers: how a norm + PReLU pair rides in the consumer's prologue.

The reference evaluates ``conv -> norm -> PReLU`` as separate modules
(conv_tasnet.py:43-49, lobe/cnn.py:62-79).  Here the producing kernel emits what the
norm needs (Welford partials for gLN/gGN), and the consuming kernel applies
``PReLU(norm(.))`` while loading its operand, so a normalised tensor is never
written to HBM.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_PRELU, PRO_AFFINE, PRO_ROWNORM, Prologue


import os

# Fused statistics finalize (the producer kernel's last CTA per item merges the partials): implemented and parity-tested,
# but measured SLOWER than the separate 5 us ps_stats_finalize launches under CUDA-graph replay (run 43, same box:
# cfg2 42.8 vs 39.9 ms, cfg4 7.5 vs 7.1 ms, cfg1 equal) - the per-tile gpu-scope fences cost the producers their L1 hits
# and the finishing CTA stalls its own pipeline.  Off by default; PS_FUSE_FINALIZE=1 turns it on.
FUSE_FINALIZE = os.environ.get("PS_FUSE_FINALIZE", "0") == "1"


def norm_kind(m: nn.Module) -> str:
    n = type(m).__name__
    if n == "GlobLN":
        return "gLN"
    if n == "ChanLN":
        return "cLN"
    if isinstance(m, nn.GroupNorm):
        if m.num_groups != 1:
            raise NotImplementedError("only GroupNorm(1, C) (gGN, lobe/norm.py:96) is on the hot path")
        return "gGN"
    if isinstance(m, nn.BatchNorm1d):
        return "bN1d"
    raise NameError("Could not interpret normalization identifier")


def needs_stats(kind: str) -> bool:
    """gLN / gGN are global over (C, T): the producer must emit Welford partials."""
    return kind in ("gLN", "gGN")


def stats_request(norm: nn.Module) -> dict:
    """Keyword arguments for the kernel that PRODUCES the tensor `norm` normalises: for gLN / gGN it must emit Welford
    partials and (fused) the folded affine, so no separate finalize launch follows."""
    kind = norm_kind(norm)
    if not FUSE_FINALIZE:
        return {"want_stats": needs_stats(kind)}
    if kind == "gLN":
        return {"want_stats": True, "fin": (norm.gamma, norm.beta, norm.eps)}
    if kind == "gGN":
        return {"want_stats": True, "fin": (norm.weight, norm.bias, norm.eps)}
    return {"want_stats": False}


def prelu_slope(m: nn.PReLU) -> torch.Tensor:
    if m.weight.numel() != 1:
        raise NotImplementedError("per-channel PReLU is not used by the reference (nn.PReLU() has one slope)")
    return m.weight


def norm_prologue(norm: nn.Module, raw: torch.Tensor, partials: Optional[torch.Tensor], slope: torch.Tensor) -> Prologue:
    """Prologue that makes a consumer see ``PReLU(norm(raw))``.  raw: [B, R, C] frames-major."""
    kind = norm_kind(norm)
    C = raw.shape[-1]
    if isinstance(partials, ops.FoldedAffine):  # the producer already finalized the statistics (stats_request)
        return Prologue(PRO_AFFINE, ACT_PRELU, partials.scale, partials.shift, C, None, slope)
    if kind == "gLN":
        scale, shift = ops.stats_finalize(partials, norm.gamma, norm.beta, norm.eps, C)
        return Prologue(PRO_AFFINE, ACT_PRELU, scale, shift, C, None, slope)
    if kind == "gGN":
        scale, shift = ops.stats_finalize(partials, norm.weight, norm.bias, norm.eps, C)
        return Prologue(PRO_AFFINE, ACT_PRELU, scale, shift, C, None, slope)
    if kind == "cLN":
        return Prologue(PRO_ROWNORM, ACT_PRELU, norm.gamma, norm.beta, 0, ops.rowstats(raw, norm.eps), slope)
    # bN1d
    if norm.training:
        raise NotImplementedError("train-mode BatchNorm couples batch items; the engine runs .eval() models only")
    scale, shift = ops.bn_fold(norm.weight, norm.bias, norm.running_mean, norm.running_var, norm.eps)
    return Prologue(PRO_AFFINE, ACT_PRELU, scale, shift, 0, None, slope)


class ParamCache:
    """Derived weight layouts (transposes, stacks, packs) keyed on the source
    parameters' storage pointer and in-place version counter, so they are rebuilt
    after ``load_state_dict`` / ``.to(device)`` and reused otherwise."""

    def __init__(self):
        self._store = {}

    def get(self, key: str, sources, build):
        sig = tuple((t.data_ptr(), t._version, t.device) for t in sources)
        hit = self._store.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = build()
        self._store[key] = (sig, val)
        return val
