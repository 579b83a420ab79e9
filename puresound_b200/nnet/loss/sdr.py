"""On-device SDR-family scoring (drop-in for the evaluation side of ``puresound.nnet.loss.sdr``: ``SDRLoss`` in eval use and
``si_snr``, reference loss/sdr.py:7-185, 263-299).  One kernel (``ps_sdr``) reads both waveforms once and
returns the score per item; the training-only options (``source_aggregated``, ``threshold``, inactive-source labels) raise.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ... import ops


def _rows(s: torch.Tensor) -> torch.Tensor:
    return s.reshape(-1, s.shape[-1]).contiguous()


def si_snr(s1: torch.Tensor, s2: torch.Tensor, eps: float = 1e-8, reduction: bool = True) -> torch.Tensor:
    """SI-SNR in dB of the estimate s1 against the reference s2, shapes [N, *, L] (reference loss/sdr.py:263-299)."""
    snr = ops.sdr(_rows(s1), _rows(s2), scaled=True, scale_dependent=False, zero_mean=True, eps=eps).view(*s1.shape[:-1], 1)
    return snr.mean() if reduction else snr


class SDRLoss(nn.Module):
    """reference: loss/sdr.py:7-185.  ``forward`` returns the NEGATIVE score like the reference (a loss)."""

    def __init__(self, scaled: bool = True, scale_dependent: bool = False, zero_mean: bool = True, source_aggregated: bool = False,
                 sdr_max: Optional[int] = None, eps: float = 1e-8, reduction: bool = True, threshold: Optional[float] = None) -> None:
        super().__init__()
        self.scaled, self.scale_dependent, self.zero_mean = scaled, scale_dependent, zero_mean
        self.source_aggregated, self.sdr_max, self.eps = source_aggregated, sdr_max, eps
        self.reduction, self.threshold = reduction, threshold
        if source_aggregated or threshold is not None:
            raise NotImplementedError("source-aggregated / thresholded SDR are training-loss options (out of scope)")

    @classmethod
    def init_mode(cls, loss_func: str = "sisnr", reduction: bool = True, threshold: Optional[float] = None):
        """Aliases of the reference (loss/sdr.py:42-101): sisnr, sdsdr, sdr, tsdr (the source-aggregated ones raise)."""
        loss_func = loss_func.lower()
        if loss_func not in ("sisnr", "sdsdr", "sdr", "tsdr", "sasdr", "sasisnr", "satsdr"):
            raise NameError
        # `loss_func in "sdsdr"` is the reference's expression (loss/sdr.py:72): a substring test, so "sdr" is scaled as well
        scaled = loss_func == "sisnr" or loss_func in "sdsdr" or loss_func == "sasisdr"
        return cls(scaled=scaled, scale_dependent=loss_func == "sdsdr", zero_mean=True,
                   source_aggregated=loss_func in ("sasdr", "sasisnr", "satsdr"), sdr_max=30 if loss_func in ("tsdr", "satsdr") else None,
                   eps=1e-8, reduction=reduction, threshold=threshold)

    @torch.no_grad()
    def forward(self, s1: torch.Tensor, s2: torch.Tensor, inactive_labels: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert s1.dim() == 2 and s2.dim() == 2, "need input shape as (batch, length)"
        if inactive_labels is not None and bool((inactive_labels == True).any()):  # noqa: E712
            raise NotImplementedError("inactive-source SDR is a training-loss option (out of scope)")
        snr = -ops.sdr(s1.contiguous(), s2.contiguous(), scaled=self.scaled, scale_dependent=self.scale_dependent, zero_mean=self.zero_mean,
                       sdr_max=self.sdr_max, eps=self.eps).view(-1, 1)
        return snr.mean() if self.reduction else snr


def align_waveform(enh_wav: torch.Tensor, ref_wav: torch.Tensor):
    """`SoTaskWrapModule._align_waveform` (base_nn.py:398-412): the reference is left-padded (aligned from the end) when it is
    shorter than the estimate, cut when longer."""
    le, lr = enh_wav.shape[-1], ref_wav.shape[-1]
    if lr < le:
        ref_wav = torch.nn.functional.pad(ref_wav, (le - lr, 0))
    elif lr > le:
        ref_wav = ref_wav[..., :le]
    return enh_wav, ref_wav
