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


class SpecAugment(nn.Module):
    """Random frequency / time band masking (reference lobe/trivial.py:307-335).

    Upstream quirk kept: the reference applies the mask in ``forward`` regardless of train / eval mode, so the speaker
    branch of ``tse_skim_v2_causal`` masks a random band of mel channels at inference too.  ``torchaudio``'s
    ``mask_along_axis`` draws ``torch.rand(1)`` twice per axis from the global CPU generator and masks ONE band for the
    whole batch; ``host_prepare`` draws the same numbers in the same order (same seed => same band as the reference) into a
    device buffer that the fill kernel reads, so the task wrapper can call it outside a captured CUDA graph."""

    def __init__(self, freq_mask_length: int, time_mask_length: int, fill_value: float) -> None:
        super().__init__()
        self.freq_mask = freq_mask_length
        self.time_mask = time_mask_length
        self.mask_value = fill_value
        self._bounds = None   # int32[4] on the device: channel band, frame band
        self._shape = None    # (channels, frames) of the masked axes last seen: the band positions depend on their lengths
        self._fresh = False
        self._stage, self._stage_i = None, 0

    @staticmethod
    def _draw(mask_param: int, axis_len: int):
        # torchaudio.functional.mask_along_axis: value = rand * mask_param; min_value = rand * (len - value)
        if mask_param < 1:
            return 0, 0
        value = torch.rand(1) * mask_param
        min_value = torch.rand(1) * (axis_len - value)
        start = int(min_value.long())
        return start, start + int(value.long())

    def host_prepare(self, device=None) -> None:
        """Draw this call's bands (host RNG) into the device buffer.  Needs the axis lengths, i.e. one earlier eager call."""
        if self._shape is None:
            return
        Cn, Tn = self._shape
        f0, f1 = self._draw(self.freq_mask, Cn)
        t0, t1 = self._draw(self.time_mask, Tn)
        # asynchronous copy from a small ring of pinned staging slots (an event per slot guards its reuse): a pageable source -
        # or a fresh pinned allocation, which is a device-wide synchronisation - would stall the serving loop
        # (inference_stream) behind the previous batch's forward
        dev = torch.device(device if device is not None else "cuda")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if self._bounds is None or self._bounds.device != dev:
            self._bounds = torch.empty(4, dtype=torch.int32, device=dev)
            self._stage = [(torch.empty(4, dtype=torch.int32).pin_memory(), torch.cuda.Event()) for _ in range(4)]
            self._stage_i = 0
        slot, ev = self._stage[self._stage_i]
        self._stage_i = (self._stage_i + 1) % len(self._stage)
        ev.synchronize()  # (returns at once unless four later calls are still queued)
        slot[0], slot[1], slot[2], slot[3] = f0, f1, t0, t1
        self._bounds.copy_(slot, non_blocking=True)
        ev.record(torch.cuda.current_stream(dev))
        self._fresh = True

    def _mask(self, x: torch.Tensor, rows: int, cols: int, channel_axis_is_rows: bool) -> torch.Tensor:
        Cn, Tn = (rows, cols) if channel_axis_is_rows else (cols, rows)
        if self.freq_mask < 1 and self.time_mask < 1:
            return x
        capturing = torch.cuda.is_current_stream_capturing()
        key = (Cn if self.freq_mask >= 1 else 0, Tn if self.time_mask >= 1 else 0)  # only the masked axes' lengths enter the draw
        if (self._shape != key or not self._fresh) and not capturing:
            self._shape = key
            self.host_prepare(x.device)
        self._fresh = False
        b = self._bounds
        if channel_axis_is_rows:  # [N, C, T] layout: the kernel's row axis is the channel axis -> swap the two bands
            b = torch.stack([b[2], b[3], b[0], b[1]])
        return ops.band_fill(x, rows, cols, b, self.mask_value)

    def forward_cl(self, x: torch.Tensor) -> torch.Tensor:
        """[N, T, C] frames-major, masked in place (the input is the mel encoder's fresh output)."""
        return self._mask(x, x.shape[1], x.shape[2], False)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[N, C, T] (reference layout) -> masked copy."""
        return self._mask(x.contiguous().clone(), x.shape[1], x.shape[2], True)


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
        pk = self._cache.get("sb_pk", [self.cond_scale.weight, self.cond_bias.weight], lambda: ops.pack_weights(w, 2 * Cn, Cn, Cn + E))
        sb, _ = ops.linear(xn, w, K=Cn, w_row_stride=Cn + E, bias_batch=eb.view(N, 2 * Cn), w_packed=pk)
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
