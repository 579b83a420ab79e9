"""Seeded synthetic inputs and weight perturbation shared by tests, bench.py and
the golden-vector generator.  Pure torch-CPU helpers; nothing in
here is on the product's compute path.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

SR = 16000
ENROLL_LEVEL = 10.0 ** (-28.0 / 20.0)  # mean-|x| the reference rescales enrollment to (src/audio.py:121-153)


def white(n: int, length: int, amp: float = 0.1, seed: int = 1234) -> torch.Tensor:
    """(i) white noise ``amp * U(-1, 1)``; amp 0.1 keeps the output clamp idle."""
    g = torch.Generator().manual_seed(seed)
    return amp * (2.0 * torch.rand(n, length, generator=g) - 1.0)


def noisy_speech(n: int, length: int, seed: int = 1234, snr_db: float = 5.0):
    """(ii) synthetic noisy speech: 5 harmonics of f0~U(90,250) Hz under a 4 Hz
    raised-cosine syllable envelope plus white noise at ``snr_db``; the mixture
    is rescaled to mean-|x| = 10^(-28/20).  Returns (mixture, clean) [n, length]."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(length, dtype=torch.float64) / SR
    f0 = 90.0 + 160.0 * torch.rand(n, 1, generator=g, dtype=torch.float64)
    phase = 2 * math.pi * torch.rand(n, 5, generator=g, dtype=torch.float64)
    env_phase = 2 * math.pi * torch.rand(n, 1, generator=g, dtype=torch.float64)
    clean = torch.zeros(n, length, dtype=torch.float64)
    for h in range(5):
        clean += torch.sin(2 * math.pi * (h + 1) * f0 * t + phase[:, h : h + 1]) / (h + 1)
    clean *= 0.5 * (1.0 - torch.cos(2 * math.pi * 4.0 * t + env_phase))
    noise = torch.randn(n, length, generator=g, dtype=torch.float64)
    ps = clean.pow(2).mean(-1, keepdim=True)
    pn = noise.pow(2).mean(-1, keepdim=True)
    noise *= torch.sqrt(ps / (pn * 10.0 ** (snr_db / 10.0)))
    mix = clean + noise
    scale = ENROLL_LEVEL / mix.abs().mean(-1, keepdim=True)
    return (mix * scale).float(), (clean * scale).float()


@torch.no_grad()
def perturb_(model: nn.Module, seed: int = 1) -> nn.Module:
    """'Perturbed' weight set of SURVEY 8d so every affine path is exercised:
    norm gains ~U(0.5,1.5), norm biases ~U(-0.5,0.5), PReLU slopes ~U(0.05,0.5),
    BatchNorm running_mean ~N(0,0.1), running_var ~U(0.5,1.5).  Conv / linear /
    LSTM weights keep their initial values.  Keyed on state-dict names in sorted
    order so the result is independent of module classes."""
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    prelu_keys = {n + ".weight" for n, m in model.named_modules() if isinstance(m, nn.PReLU)}
    norm_mods = {
        n
        for n, m in model.named_modules()
        if isinstance(m, (nn.GroupNorm, nn.LayerNorm, nn.BatchNorm1d)) or type(m).__name__ in ("GlobLN", "ChanLN")
    }
    for k in sorted(sd):
        v = sd[k]
        mod, _, leaf = k.rpartition(".")
        if k in prelu_keys:
            v.copy_(0.05 + 0.45 * torch.rand(v.shape, generator=g))
        elif mod in norm_mods:
            if leaf in ("gamma", "weight"):
                v.copy_(0.5 + torch.rand(v.shape, generator=g))
            elif leaf in ("beta", "bias"):
                v.copy_(torch.rand(v.shape, generator=g) - 0.5)
            elif leaf == "running_mean":
                v.copy_(0.1 * torch.randn(v.shape, generator=g))
            elif leaf == "running_var":
                v.copy_(0.5 + torch.rand(v.shape, generator=g))
    return model


def state_checksum(sd) -> float:
    """Order-independent fp64 checksum of a state_dict (pins seeded weights)."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k]
        if v.dtype.is_floating_point:
            tot += float(v.double().abs().sum()) + 0.5 * float(v.double().sum())
    return tot
