"""Encoders / decoders of the separator path on the B200 engine.

Drop-ins for ``puresound.nnet.lobe.encoder.FreeEncDec`` and ``ConvEncDec``
(constructor signatures, attribute names and state-dict keys preserved).

Both analysis transforms are *framed GEMMs*: the waveform is read in place as an
overlapping-row matrix (row stride = hop < K = window), so no im2col buffer is
ever materialised; the output is written frames-major ``[N, T, C]`` in exactly
the channel-cat layout the masker consumes.  Both synthesis transforms are a GEMM
to per-frame windows followed by a gather-form overlap-add with the output
constraint fused.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from ... import ops
from ...ops import ACT_NONE, ACT_RELU, GEMM_SIMT, PRO_MASK, Prologue
from .._fuse import ParamCache


_ENC_EXACT = os.environ.get("PS_ENC_EXACT", "0") == "1"


def _check_wav(x: torch.Tensor, win: int) -> None:
    if x.dim() != 2:
        raise ValueError(f"expected a waveform batch [N, L], got shape {tuple(x.shape)}")
    if x.shape[1] < win:
        # F.conv1d in the reference raises RuntimeError here (kernel larger than input)
        raise RuntimeError(f"Kernel size can't be greater than actual input size: L={x.shape[1]} < win={win}")


class FreeEncDec(nn.Module):
    """Learned Conv1d / ConvTranspose1d filterbank (reference lobe/encoder.py:16-94)."""

    def __init__(self, win_length: int = 512, laten_length: int = 512, hop_length: int = 128, output_active: bool = False):
        super().__init__()
        self.win_length = win_length
        self.hop_length = hop_length
        self.output_active = output_active
        # parameter holders under the reference's keys: encoder.weight (Nf,1,win), decoder.weight (Nf,1,win)
        self.encoder = nn.Conv1d(1, laten_length, kernel_size=win_length, stride=hop_length, bias=False)
        self.decoder = nn.ConvTranspose1d(laten_length, 1, kernel_size=win_length, stride=hop_length, bias=False)
        self._cache = ParamCache()

    # ---- engine path ----
    def encode_cl(self, wav: torch.Tensor) -> torch.Tensor:
        """wav [N, L] -> feats [N, T, Nf], T = (L - win)//hop + 1."""
        _check_wav(wav, self.win_length)
        N, L = wav.shape
        Nf, win, hop = self.encoder.out_channels, self.win_length, self.hop_length
        T = (L - win) // hop + 1
        w = self.encoder.weight.view(Nf, win)
        # tcgen05 (3xBF16, ~2^-17 per product) when the shape allows (win 32 or a multiple of 64, Nf % 32 == 0): the analysis is
        # then a pure write stream; PS_ENC_EXACT=1 keeps the exact-fp32 register filterbank kernel
        pk = None if _ENC_EXACT else self._cache.get("enc_pk", [self.encoder.weight], lambda: ops.pack_weights(w, Nf, win, win))
        y, _ = ops.gemm(wav.contiguous(), w, batch=N, rows=T, M=Nf, K=win,
                        x_batch_stride=L, x_row_stride=hop, w_row_stride=win, w_packed=pk,
                        epi_act=ACT_RELU if self.output_active else ACT_NONE)
        return y

    def decode_cl(self, feats: torch.Tensor, mask: Optional[torch.Tensor] = None, mask_act: int = ACT_NONE,
                  constraint: int = 0) -> torch.Tensor:
        """feats [N, T, Nf] (optionally multiplied by act(mask) on load) -> wav [N, (T-1)*hop + win]."""
        Nf, win = self.decoder.in_channels, self.win_length
        w_t = self._cache.get("dec_t", [self.decoder.weight], lambda: self.decoder.weight.view(Nf, win).t().contiguous())
        pro = Prologue(PRO_MASK, mask_act, x2=mask) if mask is not None else ops.NO_PRO
        # tcgen05 (3xBF16) when the shape allows: the learned synthesis filterbank has no ill-conditioned step after it
        w_pk = self._cache.get("dec_pk", [self.decoder.weight], lambda: ops.pack_weights(w_t, win, Nf, Nf))
        frames, _ = ops.linear(feats, w_t, pro=pro, w_packed=w_pk)
        return ops.ola(frames, self.hop_length, None, constraint)

    # ---- reference-layout API ----
    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[N, L] -> [N, C, T]"""
        return ops.transpose(self.encode_cl(x))

    @torch.no_grad()
    def inverse(self, x: torch.Tensor) -> torch.Tensor:
        """[N, C, T] -> [N, L']"""
        return self.decode_cl(ops.transpose(x))


class ConvSTFT(nn.Module):
    """Holder of the conv-STFT kernels under the reference's keys
    (lobe/encoder.py:275-356): ``wsin``/``wcos`` (Parameters iff trainable) and the
    iSTFT buffers ``kernel_sin_inv``/``kernel_cos_inv``/``window_mask``."""

    def __init__(self, window: torch.Tensor, n_fft: int, hop_length: int, iSTFT: bool, trainable: bool):
        super().__init__()
        self.n_fft, self.stride, self.trainable, self.iSTFT = n_fft, hop_length, trainable, iSTFT
        if len(window) != n_fft:
            raise TypeError("only support window length == n_fft")
        # float64 Fourier kernels cast to fp32, bins 0..n_fft/2 (lobe/stft.py:91-100, freq_scale='no')
        s = np.arange(0, n_fft, 1.0)
        bins = n_fft // 2 + 1
        ksin = np.empty((bins, 1, n_fft))
        kcos = np.empty((bins, 1, n_fft))
        for k in range(bins):
            ksin[k, 0, :] = np.sin(2 * np.pi * k * s / n_fft)
            kcos[k, 0, :] = np.cos(2 * np.pi * k * s / n_fft)
        ksin = torch.tensor(ksin.astype(np.float32), dtype=torch.float)
        kcos = torch.tensor(kcos.astype(np.float32), dtype=torch.float)
        if iSTFT:
            self.register_buffer("kernel_sin_inv", torch.cat((ksin, -ksin[1:-1].flip(0)), 0).unsqueeze(-1))
            self.register_buffer("kernel_cos_inv", torch.cat((kcos, kcos[1:-1].flip(0)), 0).unsqueeze(-1))
        wsin, wcos = ksin * window, kcos * window
        if trainable:
            self.wsin = nn.Parameter(wsin)
            self.wcos = nn.Parameter(wcos)
        else:
            self.register_buffer("wsin", wsin)
            self.register_buffer("wcos", wcos)
        self.register_buffer("window_mask", window.unsqueeze(0).unsqueeze(-1))


def _mel_scale(hz: np.ndarray) -> np.ndarray:
    """Slaney mel scale (lobe/stft.py:127-158): linear below 1 kHz (200/3 Hz per mel), logarithmic above (27 mels per factor 6.4)."""
    hz = np.asarray(hz, dtype=np.float64)
    lin = hz / (200.0 / 3)
    brk = 1000.0 / (200.0 / 3)
    with np.errstate(divide="ignore", invalid="ignore"):
        log = brk + np.log(hz / 1000.0) / (np.log(6.4) / 27.0)
    return np.where(hz >= 1000.0, log, lin)


def _mel_scale_inv(mel: np.ndarray) -> np.ndarray:
    """Inverse of _mel_scale (lobe/stft.py:161-190)."""
    mel = np.asarray(mel, dtype=np.float64)
    brk = 1000.0 / (200.0 / 3)
    return np.where(mel >= brk, 1000.0 * np.exp((np.log(6.4) / 27.0) * (mel - brk)), (200.0 / 3) * mel)


def mel_filterbank(sr: int, n_fft: int, n_banks: int = 128, fmin: float = 0.0, fmax: Optional[float] = None) -> torch.Tensor:
    """Triangular, area-normalised (Slaney) mel filters [n_banks, n_fft//2 + 1], fp32 — the values of the reference's
    ``mel_filterbank`` (lobe/stft.py:237-293): band edges uniform on the mel scale between fmin and fmax (default
    Nyquist), each triangle = max(0, min(rising ramp, falling ramp)) over the FFT bin centres, scaled by 2 / band width."""
    fmax = float(sr / 2) if fmax is None else fmax
    bins = np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)
    edges = _mel_scale_inv(np.linspace(_mel_scale(fmin), _mel_scale(fmax), n_banks + 2))
    width = np.diff(edges)
    ramps = np.subtract.outer(edges, bins)                      # [n_banks + 2, bins]: edge - bin centre
    rising = -ramps[:-2] / width[:-1, None]
    falling = ramps[2:] / width[1:, None]
    tri = np.maximum(0, np.minimum(rising, falling)).astype(np.float32)
    tri *= (2.0 / (edges[2:n_banks + 2] - edges[:n_banks]))[:, None]
    if not np.all((edges[:-2] == 0) | (tri.max(axis=1) > 0)):
        raise ValueError("Empty filters detected in mel frequency basis.")
    return torch.from_numpy(tri)


class ConvMelSpectrogram(ConvSTFT):
    """Holder of the mel front-end's tensors under the reference's keys (lobe/encoder.py:459-507): the conv-STFT kernels
    plus ``filterbank`` [bins, n_banks] and its pseudo-inverse ``inv_filterbank`` (Parameters iff trainable)."""

    def __init__(self, window: torch.Tensor, n_fft: int, hop_length: int, trainable: bool, n_banks: int):
        super().__init__(window, n_fft, hop_length, False, trainable)
        fb = mel_filterbank(sr=16000, n_fft=n_fft, n_banks=n_banks).permute(1, 0)  # the reference fixes sr = 16 kHz here (:494-496)
        inv = torch.pinverse(fb)
        if trainable:
            self.filterbank = nn.Parameter(fb)
            self.inv_filterbank = nn.Parameter(inv)
        else:
            self.register_buffer("filterbank", fb)
            self.register_buffer("inv_filterbank", inv)


class FbankEnc(nn.Module):
    """Mel-spectrogram speaker front-end (reference lobe/encoder.py:186-272; used as ``encoder_spk`` by
    ``tse_skim_v2_causal``, egs/tse/model.py:519-521): power spectrum of the conv-STFT times the mel filterbank.

    Engine path: the framed analysis GEMM (cos | -sin, as ConvEncDec) gives X [N, T, 2F]; the mel projection is ONE more
    GEMM over K = 2F with the square taken on load (the mask prologue with the operand as its own mask) against the
    filterbank stacked twice - mel[m] = sum_f fb[f, m] (re_f^2 + im_f^2) - so the power spectrum is never written."""

    def __init__(
        self,
        fft_length: int = 512,
        win_type: str = "hann",
        win_length: int = 512,
        freq_bins: int = None,
        hop_length: int = 128,
        freq_scale: str = "no",
        fmin: int = 0,
        fmax: int = 8000,
        sr: int = 16000,
        trainable: bool = True,
        output_format: str = "Magnitude",
        n_banks=80,
    ):
        super().__init__()
        if freq_scale != "no" or freq_bins is not None:
            raise NotImplementedError("only the linear full-band STFT (freq_scale='no') is on the separator path")
        if output_format.lower() != "magnitude":
            raise NotImplementedError("only output_format='Magnitude' feeds a speaker net (MagPhase / inverse are synthesis-side)")
        self.n_fft, self.win_length, self.freq_bins, self.hop_length = fft_length, win_length, freq_bins, hop_length
        self.freq_scale, self.iSTFT, self.fmin, self.fmax, self.sr = freq_scale, False, fmin, fmax, sr
        self.trainable, self.output_format, self.n_banks = trainable, output_format, n_banks
        if win_type.lower() != "hann":
            raise NotImplementedError("window type not support")
        self.window = torch.hann_window(win_length)
        self.encoder = ConvMelSpectrogram(self.window, fft_length, hop_length, trainable, n_banks)
        self._cache = ParamCache()

    def encode_cl(self, wav: torch.Tensor, exact: bool = True) -> torch.Tensor:
        """wav [N, L] -> mel power spectrum [N, T, n_banks]."""
        _check_wav(wav, self.n_fft)
        e = self.encoder
        F2 = 2 * e.wcos.shape[0]
        Mp = (F2 + 31) // 32 * 32  # 514 -> 544 zero rows: a channel count the tcgen05 kernel takes (their mel weights are zero too)
        w = self._cache.get("ana", [e.wsin, e.wcos], lambda: torch.cat(
            [e.wcos[:, 0, :], -e.wsin[:, 0, :], e.wcos.new_zeros(Mp - F2, self.n_fft)], 0).contiguous())
        N, L = wav.shape
        T = (L - self.n_fft) // self.hop_length + 1
        pk = None
        if not exact:
            pk = self._cache.get("ana_pk", [e.wsin, e.wcos], lambda: ops.pack_weights(w, w.shape[0], self.n_fft, self.n_fft))
        X, _ = ops.gemm(wav.contiguous(), w, batch=N, rows=T, M=w.shape[0], K=self.n_fft, x_batch_stride=L,
                        x_row_stride=self.hop_length, w_row_stride=self.n_fft, w_packed=pk,
                        backend=ops.GEMM_AUTO if pk is not None else GEMM_SIMT)
        fb2 = self._cache.get("fb2", [e.filterbank], lambda: torch.cat(
            [e.filterbank, e.filterbank, e.filterbank.new_zeros(Mp - F2, e.filterbank.shape[1])], 0).t().contiguous())  # [n_banks, Mp]
        # trainable: the reference adds 1e-8 to every power bin before the projection (:530) = a per-band constant
        bias = self._cache.get("fb_bias", [e.filterbank], lambda: (1e-8 * e.filterbank.sum(0)).contiguous()) if self.trainable else None
        mel, _ = ops.linear(X, fb2, pro=Prologue(PRO_MASK, ACT_NONE, x2=X), bias=bias, backend=GEMM_SIMT)
        return mel

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[N, L] -> [N, n_banks, T]"""
        return ops.transpose(self.encode_cl(x))

    def inverse(self, magphase: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("Inverse only support magphase input")


class ConvEncDec(nn.Module):
    """Conv-STFT encoder / conv-iSTFT decoder (reference lobe/encoder.py:97-183)."""

    def __init__(
        self,
        fft_length: int = 512,
        win_type: str = "hann",
        win_length: int = 512,
        freq_bins: int = None,
        hop_length: int = 128,
        freq_scale: str = "no",
        iSTFT: bool = True,
        fmin: int = 0,
        fmax: int = 8000,
        sr: int = 16000,
        trainable: bool = True,
        output_format: str = "Complex",
    ):
        super().__init__()
        if freq_scale != "no" or freq_bins is not None:
            raise NotImplementedError("only the linear full-band STFT (freq_scale='no') is on the separator path")
        if output_format != "Complex":
            raise NotImplementedError("only output_format='Complex' feeds the maskers")
        self.n_fft, self.win_length, self.freq_bins, self.hop_length = fft_length, win_length, freq_bins, hop_length
        self.freq_scale, self.iSTFT, self.fmin, self.fmax, self.sr = freq_scale, iSTFT, fmin, fmax, sr
        self.trainable, self.output_format = trainable, output_format
        if win_type.lower() != "hann":
            raise NotImplementedError("window type not support")
        self.window = torch.hann_window(win_length)
        self.encoder = ConvSTFT(self.window, fft_length, hop_length, iSTFT, trainable)
        self._cache = ParamCache()
        self._wsum = {}

    @property
    def bins(self) -> int:
        return self.n_fft // 2 + 1

    # ---- engine path ----
    def encode_cl(self, wav: torch.Tensor, drop_first_bin: bool, exact: bool = True) -> torch.Tensor:
        """wav [N, L] -> [N, T, 2F'] = cat(Re[s:], Im[s:]) on channels (base_nn.py:337-345 layout), Im = -conv(wsin).

        exact=True keeps the analysis GEMM in true fp32; exact=False (what the task wrapper uses, see base_nn._STFT_EXACT)
        lets it run on the tensor cores (3xBF16).  The synthesis side (decode_cl) is always exact fp32: its
        window-sum-square division amplifies rounding ~2.6e4x at the first/last hop."""
        _check_wav(wav, self.n_fft)
        e = self.encoder
        s = 1 if drop_first_bin else 0
        w = self._cache.get(f"ana{s}", [e.wsin, e.wcos], lambda: torch.cat([e.wcos[s:, 0, :], -e.wsin[s:, 0, :]], 0).contiguous())
        N, L = wav.shape
        T = (L - self.n_fft) // self.hop_length + 1
        pk = None
        if not exact:
            pk = self._cache.get(f"ana_pk{s}", [e.wsin, e.wcos], lambda: ops.pack_weights(w, w.shape[0], self.n_fft, self.n_fft))
        y, _ = ops.gemm(wav.contiguous(), w, batch=N, rows=T, M=w.shape[0], K=self.n_fft, x_batch_stride=L,
                        x_row_stride=self.hop_length, w_row_stride=self.n_fft, w_packed=pk,
                        backend=ops.GEMM_AUTO if pk is not None else GEMM_SIMT)
        return y

    def _synthesis_weight(self, s: int) -> torch.Tensor:
        """[n_fft (sample k), 2F'] such that frame[k] = sum_f Re[f] Wre[k,f] + Im[f] Wim[k,f] equals the reference's
        Hermitian-extended conv2d pair, window and 1/n_fft included (lobe/encoder.py:419-438, lobe/stft.py:118-125)."""
        e = self.encoder
        F = self.bins

        def build():
            kc = e.kernel_cos_inv[:F, 0, :, 0].double()  # [F, n_fft]
            ks = e.kernel_sin_inv[:F, 0, :, 0].double()
            coef = torch.full((F, 1), 2.0, dtype=torch.float64, device=kc.device)
            coef[0] = 1.0
            coef[F - 1] = 1.0
            win = e.window_mask.flatten().double() / self.n_fft
            wre = (coef * kc * win)[s:]
            wim = (-coef * ks * win)[s:]
            return torch.cat([wre, wim], 0).t().contiguous().float()

        return self._cache.get(f"syn{s}", [e.kernel_cos_inv, e.kernel_sin_inv, e.window_mask], build)

    def _window_sumsquare(self, T: int, device) -> torch.Tensor:
        """sum_t window^2[j - t*hop] (lobe/stft.py:109-115); constant per frame count, cached."""
        key = (T, str(device), self.encoder.window_mask._version)
        if key not in self._wsum:
            w2 = (self.encoder.window_mask.detach().flatten().cpu() ** 2)
            out_len = self.n_fft + self.hop_length * (T - 1)
            stack = w2.unsqueeze(-1).repeat(1, T).unsqueeze(0)
            ws = torch.nn.functional.fold(stack, (1, out_len), kernel_size=(1, self.n_fft), stride=self.hop_length).flatten()
            # insert, never replace: a captured CUDA graph bakes in the device pointer of the entry it was captured with
            # (SoTaskWrapModule keeps several shape slots alive), so an entry must stay allocated while the module lives
            if len(self._wsum) >= 64:
                self._wsum.pop(next(iter(self._wsum)))
            self._wsum[key] = ws.to(device)
        return self._wsum[key]

    def decode_cl(self, feats: torch.Tensor, drop_first_bin: bool, constraint: int = 0) -> torch.Tensor:
        """[N, T, 2F'] channel-cat spectrum -> wav [N, n_fft + hop*(T-1)].  The synthesis GEMM and the
        overlap-add stay in true fp32: the division by the window sum-square amplifies rounding ~2.6e4x at
        the first/last hop."""
        if not self.iSTFT:
            raise NameError("Please activate the iSTFT module by setting `iSTFT=True` if you want to use `inverse`")
        s = 1 if drop_first_bin else 0
        w = self._synthesis_weight(s)
        frames, _ = ops.linear(feats, w, backend=GEMM_SIMT)
        return ops.ola(frames, self.hop_length, self._window_sumsquare(feats.shape[1], feats.device), constraint)

    # ---- reference-layout API ----
    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[N, L] -> [N, F, T, 2] (real, imag)"""
        y = ops.transpose(self.encode_cl(x, False))  # [N, 2F, T]
        N, _, T = y.shape
        return y.view(N, 2, self.bins, T).permute(0, 2, 3, 1).contiguous()

    @torch.no_grad()
    def inverse(self, x: torch.Tensor) -> torch.Tensor:
        """[N, F, T, 2] -> [N, L']"""
        assert x.dim() == 4, "Inverse iSTFT only works for complex number, shape (batch, freq_bins, timesteps, 2)"
        N, F, T, _ = x.shape
        cl = x.permute(0, 2, 3, 1).reshape(N, T, 2 * F).contiguous()
        return self.decode_cl(cl, False)
