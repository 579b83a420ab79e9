"""Turn a task-wrapper module tree into the plain-dict config the oracle takes.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Works on class *names* and on the
attribute names the reference uses (base_nn.py:247-258, conv_tasnet.py:256-268,
dprnn.py:43-54, encoder.py:36-38,125-135), so the same function describes both
the reference's modules (when generating golden vectors in the authoring
container) and this repo's drop-in modules (in the GPU parity tests).
"""
from __future__ import annotations

import torch.nn as nn


def _norm_kind(m: nn.Module) -> str:
    n = type(m).__name__
    return {"GlobLN": "gLN", "ChanLN": "cLN", "GroupNorm": "gGN", "BatchNorm1d": "bN1d"}[n]


def describe_tcn(m: nn.Module) -> dict:
    dw = m.dconv[0].depthwise[0]
    return {
        "type": "TCN",
        "in_channels": m.out_conv.out_channels,
        "hid_channels": m.out_conv.in_channels,
        "emb_dim": m.in_conv[0].in_channels - m.out_conv.out_channels,
        "kernel": dw.kernel_size[0],
        "dilation": dw.dilation[0],
        "causal": bool(m.dconv[0].causal),
        "tcn_norm": _norm_kind(m.in_conv[1]),
        "dconv_norm": _norm_kind(m.dconv[0].depthwise[1]),
    }


def describe_gated_tcn(m: nn.Module) -> dict:
    c = m.left_conv[0]
    return {
        "type": "GatedTCN",
        "in_channels": m.out_conv.out_channels,
        "hid_channels": m.out_conv.in_channels,
        "emb_dim": (m.cond_scale.in_channels if m.use_film else m.right_conv[0].in_channels - m.out_conv.in_channels),
        "kernel": c.kernel_size[0],
        "dilation": c.dilation[0],
        "causal": bool(m.causal),
        "tcn_norm": _norm_kind(m.left_conv[1]),
        "use_film": bool(m.use_film),
    }


def describe_encoder(m: nn.Module) -> dict:
    n = type(m).__name__
    if n == "FreeEncDec":
        return {
            "type": "FreeEncDec",
            "win_length": m.win_length,
            "hop_length": m.hop_length,
            "laten_length": m.encoder.out_channels,
            "output_active": bool(m.output_active),
        }
    if n == "ConvEncDec":
        return {"type": "ConvEncDec", "fft_length": m.n_fft, "win_length": m.win_length, "hop_length": m.hop_length}
    if n == "FbankEnc":
        return {"type": "FbankEnc", "fft_length": m.n_fft, "hop_length": m.hop_length, "n_banks": m.n_banks, "trainable": bool(m.trainable)}
    raise NotImplementedError(n)


def describe_masker(m: nn.Module) -> dict:
    n = type(m).__name__
    if n in ("ConvTasNet", "StreamingConvTasNet"):
        return {"type": "ConvTasNet", **m.get_args}
    if n in ("UnetTcn", "DPCRN", "DPARN"):
        a = {k: (list(v) if isinstance(v, tuple) else v) for k, v in m.get_args.items()}
        if n == "DPARN":  # the reference's get_args omits nhead (dparn.py:226-246): read it off the attention module
            a["nhead"] = m.dprnn_block1.intra_atten1.self_atten.atten.num_heads
        if a["input_type"].lower() == "ri":  # get_args reports the channel list AFTER the constructor doubled entry 0 (unet.py:91-93)
            a["channels"] = [a["channels"][0] // 2] + list(a["channels"][1:])
        return {"type": n, **a}
    if n == "DPRNN":
        return {
            "type": "DPRNN",
            "input_size": m.input_size,
            "hidden_size": m.hidden_size,
            "n_blocks": m.n_blocks,
            "seg_size": m.seg_size,
            "seg_overlap": bool(m.seg_overlap),
            "causal": not m.bi_direct,
            "embed_dim": m.embed_dim,
            "embed_norm": bool(m.embed_norm),
            "block_with_embed": None if m.block_with_embed is None else [bool(b) for b in m.block_with_embed],
            "embedding_free_tse": bool(m.embedding_free_tse),
        }
    if n in ("SkiM", "StreamingSkiM"):
        fusion = None
        if m.embed_dim != 0:
            kinds = {type(f).__name__ for f in m.seg_input_fusion if f is not None}
            if kinds - {"FiLM"}:
                raise NotImplementedError(f"SkiM embedding fusion {kinds}")
            fusion = "FiLM"
        return {
            "type": "SkiM",
            "input_size": m.seg_lstm[0].input_size,
            "hidden_size": m.hidden_size,
            "n_blocks": m.n_blocks,
            "seg_size": m.seg_size,
            "seg_overlap": bool(m.seg_overlap),
            "causal": bool(m.causal),
            "embed_dim": m.embed_dim,
            "embed_norm": bool(m.embed_norm),
            "embed_fusion": fusion,
            "block_with_embed": None if m.block_with_embed is None else [bool(b) for b in m.block_with_embed],
        }
    raise NotImplementedError(n)


def describe_speaker_net(m) -> list:
    layers = list(m) if isinstance(m, (nn.ModuleList, nn.Sequential)) else [m]
    out = []
    for l in layers:
        n = type(l).__name__
        if n == "Magnitude":
            out.append({"type": "Magnitude", "drop_first": bool(l.drop_first), "log1p": bool(l.log1p)})
        elif n == "TCN":
            out.append(describe_tcn(l))
        elif n == "GatedTCN":
            out.append(describe_gated_tcn(l))
        elif n == "AttentiveStatisticsPooling":
            out.append({"type": n, "channels": l.conv.out_channels, "attention_channels": l.conv.in_channels})
        elif n == "SpecAugment":
            out.append({"type": n, "freq_mask": l.freq_mask, "time_mask": l.time_mask, "mask_value": l.mask_value})
        elif n == "SingleRNN":
            out.append({"type": n, "bidirectional": l.num_direction == 2})
        elif n == "Conv1d":
            assert l.kernel_size == (1,)
            out.append({"type": "Conv1d", "bias": l.bias is not None})
        else:
            raise NotImplementedError(n)
    return out


def describe(model: nn.Module) -> dict:
    """Config of a SoTaskWrapModule-shaped module."""
    return {
        "encoder": describe_encoder(model.encoder),
        "encoder_spk": None if model.encoder_spk is None else describe_encoder(model.encoder_spk),
        "masker": describe_masker(model.masker),
        "speaker_net": None if model.speaker_net is None else describe_speaker_net(model.speaker_net),
        "embedding_free_tse": bool(model.embedding_free_tse),
        "f_type": model.f_type,
        "mask_type": model.mask_type,
        "mask_constraint": model.mask_constraint,
        "output_constraint": model.output_constraint,
        "drop_first_bin": bool(model.drop_first_bin),
    }
