"""Model registry: the reference's recipe constructors (egs/tse/model.py:89-182,
608-637) and the BASELINE.json configurations, built from this
package's drop-in modules.  The constructor arguments *are* the configuration —
the reference has no other config system on this path.
"""
from __future__ import annotations

from typing import Optional

import torch.nn as nn

from .nnet.base_nn import SoTaskWrapModule
from .nnet.conv_tasnet import TCN, ConvTasNet, GatedTCN
from .nnet.dparn import DPARN
from .nnet.dpcrn import DPCRN
from .nnet.dprnn import DPRNN
from .nnet.lobe.encoder import ConvEncDec, FbankEnc, FreeEncDec
from .nnet.lobe.pooling import AttentiveStatisticsPooling
from .nnet.lobe.rnn import SingleRNN
from .nnet.lobe.trivial import Magnitude, SpecAugment
from .nnet.skim import SkiM
from .nnet.unet import UnetTcn


def _td_speaker_net():
    return nn.ModuleList(
        [TCN(512, 256, 3, dilation=2 ** i, causal=False, tcn_norm="gLN", dconv_norm="gGN") for i in range(5)]
        + [AttentiveStatisticsPooling(512, 128), nn.Conv1d(512 * 2, 192, 1, bias=False)]
    )


def init_model(name: str, sig_loss: Optional[nn.Module] = None, cls_loss: Optional[nn.Module] = None, **kwargs):
    """Same names and argument meaning as the reference's ``init_model`` (egs/tse/model.py:89-94);
    unknown names raise NameError like the reference (:639-640)."""
    if name in ("td_tse_conv_tasnet_v0", "td_tse_conv_tasnet_v0_causal"):
        causal = name.endswith("_causal")
        norm = "bN1d" if causal else "gLN"
        dnorm = "bN1d" if causal else "gGN"
        return SoTaskWrapModule(
            encoder=FreeEncDec(win_length=32, hop_length=16, laten_length=512),
            masker=ConvTasNet(512, 192, True, tcn_kernel=3, tcn_dim=256, repeat_tcn=3, tcn_dilated_basic=2, per_tcn_stack=8,
                              tcn_with_embed=[1, 0, 0, 0, 0, 0, 0, 0], tcn_norm=norm, dconv_norm=dnorm, causal=causal, tcn_layer="normal"),
            speaker_net=_td_speaker_net(),
            loss_func_wav=sig_loss, loss_func_spk=cls_loss, mask_constraint="ReLU", **kwargs)
    if name == "veve_dprnn_v0_causal":
        return SoTaskWrapModule(
            encoder=FreeEncDec(win_length=32, hop_length=16, laten_length=128, output_active=True),
            masker=DPRNN(input_size=128, hidden_size=64, output_size=128, n_blocks=6, seg_size=20, seg_overlap=False, causal=True,
                         embed_dim=0, embed_norm=False, block_with_embed=(False,) * 6, embedding_free_tse=True),
            speaker_net=None, loss_func_wav=sig_loss, loss_func_spk=cls_loss, mask_constraint="ReLU", embedding_free_tse=True, **kwargs)
    if name in ("tse_unet_tcn_v0", "tse_unet_tcn_v0_causal", "tse_unet_tcn_v1"):
        # egs/tse/model.py:184-369: STFT 512/128, 6-layer U-Net shell (kernel 5x2, frequency stride 2) around 3 x 5 GatedTCN
        # blocks (embedding concatenated - v1: FiLM - in the first block of each repeat), GatedTCN speaker net
        causal = name.endswith("_causal")
        return SoTaskWrapModule(
            encoder=ConvEncDec(fft_length=512, win_type="hann", win_length=512, hop_length=128, trainable=True, output_format="Complex"),
            masker=UnetTcn(
                embed_dim=192, embed_norm=True, input_type="RI", input_dim=512, activation_type="PReLU",
                norm_type="bN2d" if causal else "gLN", channels=(1, 32, 64, 128, 128, 128, 128), transpose_t_size=2, transpose_delay=True,
                skip_conv=False, kernel_t=(2,) * 6, kernel_f=(5,) * 6, stride_t=(1,) * 6, stride_f=(2,) * 6, dilation_t=(1,) * 6,
                dilation_f=(1,) * 6, delay=(0,) * 6, tcn_layer="gated", tcn_kernel=3, tcn_dim=256, tcn_dilated_basic=2, per_tcn_stack=5,
                repeat_tcn=3, tcn_with_embed=[1, 0, 0, 0, 0], tcn_norm="bN1d" if causal else "gLN", dconv_norm="bN1d" if causal else "gGN",
                causal=causal, **({"tcn_use_film": True} if name.endswith("_v1") else {})),
            speaker_net=nn.ModuleList(
                [Magnitude(drop_first=False)] + [GatedTCN(256, 128, 3, dilation=2 ** i, causal=False, tcn_norm="gLN") for i in range(5)]
                + [AttentiveStatisticsPooling(256, 128), nn.Conv1d(256 * 2, 192, 1, bias=False)]),
            loss_func_wav=sig_loss, loss_func_spk=cls_loss, mask_constraint="linear", drop_first_bin=True, **kwargs)
    if name in ("ns_dpcrn_v0", "ns_dpcrn_v0_causal"):
        # egs/ns/model.py:38-126 (noise suppression: complex mask on the STFT, no speaker branch); the non-causal variant
        # only differs in which frame the up-path trim drops (transpose_delay)
        return SoTaskWrapModule(
            encoder=ConvEncDec(fft_length=512, win_type="hann", win_length=512, hop_length=128, trainable=True, output_format="Complex"),
            masker=DPCRN(input_type="RI", input_dim=512, activation_type="PReLU", norm_type="bN2d", dropout=0.1,
                         channels=(1, 32, 32, 32, 64, 128), transpose_t_size=2, transpose_delay=not name.endswith("_causal"), skip_conv=False,
                         kernel_t=(2,) * 5, kernel_f=(5, 3, 3, 3, 3), stride_t=(1,) * 5, stride_f=(2, 2, 1, 1, 1), dilation_t=(1,) * 5,
                         dilation_f=(1,) * 5, delay=(0,) * 5, rnn_hidden=128),
            speaker_net=None, loss_func_wav=sig_loss, loss_func_spk=None, drop_first_bin=True, mask_constraint="linear",
            f_type="Complex", mask_type="Complex", **kwargs)
    if name in ("ns_dparn_v0", "ns_dparn_v0_causal"):
        # egs/ns/model.py:128-216: DPCRN with the intra-chunk LSTM replaced by two 8-head transformer encoder layers
        return SoTaskWrapModule(
            encoder=ConvEncDec(fft_length=512, win_type="hann", win_length=512, hop_length=128, trainable=True, output_format="Complex"),
            masker=DPARN(input_type="RI", input_dim=512, activation_type="PReLU", norm_type="bN2d", dropout=0.1,
                         channels=(1, 32, 32, 32, 64, 128), transpose_t_size=2, transpose_delay=not name.endswith("_causal"), skip_conv=False,
                         kernel_t=(2,) * 5, kernel_f=(5, 3, 3, 3, 3), stride_t=(1,) * 5, stride_f=(2, 2, 1, 1, 1), dilation_t=(1,) * 5,
                         dilation_f=(1,) * 5, delay=(0,) * 5, rnn_hidden=128, nhead=8),
            speaker_net=None, loss_func_wav=sig_loss, loss_func_spk=None, drop_first_bin=True, mask_constraint="linear",
            f_type="Complex", mask_type="Complex", **kwargs)
    if name in ("tse_skim_v0", "tse_skim_v0_causal"):
        # egs/tse/model.py:371-463 (the causal one is the reference's demo model)
        causal = name.endswith("_causal")
        return SoTaskWrapModule(
            encoder=FreeEncDec(win_length=32, hop_length=16, laten_length=128, output_active=True),
            masker=SkiM(input_size=128, hidden_size=256, output_size=128, n_blocks=4, seg_size=150, seg_overlap=False, causal=causal,
                        embed_dim=192, embed_norm=True, block_with_embed=[1, 1, 1, 1], embed_fusion="FiLM"),
            speaker_net=nn.ModuleList(
                [TCN(128, 256, 3, dilation=2 ** i, causal=False, tcn_norm="gLN", dconv_norm="gGN") for i in range(5)]
                + [AttentiveStatisticsPooling(128, 128), nn.Conv1d(128 * 2, 192, 1, bias=False)]),
            loss_func_wav=sig_loss, loss_func_spk=cls_loss, mask_constraint="ReLU", **kwargs)
    if name == "tse_skim_v0_causal_vad":
        # egs/tse/model.py:560-606: the small causal SkiM (hidden 64, 2 blocks) with the sigmoid output constraint
        return SoTaskWrapModule(
            encoder=FreeEncDec(win_length=32, hop_length=16, laten_length=128, output_active=True),
            masker=SkiM(input_size=128, hidden_size=64, output_size=128, n_blocks=2, seg_size=150, seg_overlap=False, causal=True,
                        embed_dim=192, embed_norm=True, block_with_embed=[1, 1], embed_fusion="FiLM"),
            speaker_net=nn.ModuleList(
                [TCN(128, 256, 3, dilation=2 ** i, causal=False, tcn_norm="gLN", dconv_norm="gGN") for i in range(5)]
                + [AttentiveStatisticsPooling(128, 128), nn.Conv1d(128 * 2, 192, 1, bias=False)]),
            loss_func_wav=sig_loss, loss_func_spk=cls_loss, mask_constraint="ReLU", output_constraint="Sigmoid", **kwargs)
    if name in ("tse_skim_v1_causal", "tse_skim_v2_causal"):
        # egs/tse/model.py:465-549: the causal SkiM masker with (v1) a bidirectional-LSTM speaker net on the learned encoder's
        # features, or (v2) a mel front-end (FbankEnc, 80 bands, fixed filters) + SpecAugment + five TCN blocks
        # (constructed in the reference's keyword order - encoder, encoder_spk, masker, speaker_net - so a seeded init matches)
        v2 = name == "tse_skim_v2_causal"
        encoder = FreeEncDec(win_length=32, hop_length=16, laten_length=128, output_active=True)
        spk_enc = FbankEnc(trainable=False, output_format="Magnitude", n_banks=80) if v2 else None
        masker = SkiM(input_size=128, hidden_size=256, output_size=128, n_blocks=4, seg_size=150, seg_overlap=False, causal=True,
                      embed_dim=192, embed_norm=True, block_with_embed=[1, 1, 1, 1], embed_fusion="FiLM")
        if v2:
            spk = ([SpecAugment(freq_mask_length=10, time_mask_length=0, fill_value=0.0)]
                   + [TCN(80, 256, 3, dilation=2 ** i, causal=False, tcn_norm="gLN", dconv_norm="gGN") for i in range(5)]
                   + [AttentiveStatisticsPooling(80, 128), nn.Conv1d(80 * 2, 192, 1, bias=False)])
        else:
            spk = [SingleRNN(rnn_type="LSTM", input_size=128, hidden_size=192, bidirectional=True, dropout=0.05),
                   AttentiveStatisticsPooling(128, 128), nn.Conv1d(128 * 2, 192, 1, bias=False)]
        return SoTaskWrapModule(
            encoder=encoder, encoder_spk=spk_enc, masker=masker, speaker_net=nn.ModuleList(spk),
            loss_func_wav=sig_loss, loss_func_spk=cls_loss, mask_constraint="ReLU", **kwargs)
    raise NameError


def baseline_config(name: str, verbose: bool = False) -> SoTaskWrapModule:
    """The five BASELINE.json configurations as SURVEY.md 8d maps them onto reference constructors."""
    if name in ("cfg1", "cfg2"):
        return SoTaskWrapModule(
            FreeEncDec(32, 512, 16),
            ConvTasNet(512, 0, False, tcn_kernel=3, tcn_dim=512, repeat_tcn=3, tcn_dilated_basic=2, per_tcn_stack=8,
                       tcn_with_embed=[0] * 8, tcn_norm="gLN", dconv_norm="gGN", causal=False, tcn_layer="normal"),
            mask_constraint="ReLU", verbose=verbose)
    if name == "cfg1b":
        # the reading of BASELINE's (N=512, B=128, H=512) that honours B - a 128-wide residual stream
        # (reference constructor conv_tasnet.py:239-254 with input_dim=128) around 512-wide blocks; 9,583,688 parameters
        return SoTaskWrapModule(
            FreeEncDec(32, 128, 16),
            ConvTasNet(128, 0, False, tcn_kernel=3, tcn_dim=512, repeat_tcn=3, tcn_dilated_basic=2, per_tcn_stack=8,
                       tcn_with_embed=[0] * 8, tcn_norm="gLN", dconv_norm="gGN", causal=False, tcn_layer="normal"),
            mask_constraint="ReLU", verbose=verbose)
    if name == "cfg3":
        return SoTaskWrapModule(
            FreeEncDec(32, 128, 16, output_active=True),
            DPRNN(128, 128, 128, n_blocks=6, seg_size=100, seg_overlap=True, causal=False),
            mask_constraint="ReLU", verbose=verbose)
    if name == "cfg4":
        return SoTaskWrapModule(
            ConvEncDec(512, "hann", 512, hop_length=128, trainable=True, output_format="Complex"),
            ConvTasNet(512, 192, True, tcn_dim=256, repeat_tcn=3, per_tcn_stack=8, tcn_with_embed=[1, 0, 0, 0, 0, 0, 0, 0]),
            speaker_net=nn.ModuleList([Magnitude(drop_first=False)] + [TCN(256, 256, 3, dilation=2 ** i) for i in range(5)]
                                      + [AttentiveStatisticsPooling(256, 128), nn.Conv1d(512, 192, 1, bias=False)]),
            mask_constraint="linear", drop_first_bin=True, verbose=verbose)
    if name == "cfg4_gated":
        # cfg4 with the GatedTCN blocks of the reference's STFT-domain TSE recipes (tse_unet_tcn_v0, egs/tse/model.py:184-243:
        # gated masker blocks tcn_dim 256 with the embedding concatenated in block 0 of each repeat, and their speaker net
        # of five GatedTCN(256, 128) blocks) - the recipes' U-Net shell itself is a later row
        return SoTaskWrapModule(
            ConvEncDec(512, "hann", 512, hop_length=128, trainable=True, output_format="Complex"),
            ConvTasNet(512, 192, True, tcn_layer="gated", tcn_dim=256, repeat_tcn=3, per_tcn_stack=5, tcn_with_embed=[1, 0, 0, 0, 0]),
            speaker_net=nn.ModuleList([Magnitude(drop_first=False)]
                                      + [GatedTCN(256, 128, 3, dilation=2 ** i, causal=False, tcn_norm="gLN") for i in range(5)]
                                      + [AttentiveStatisticsPooling(256, 128), nn.Conv1d(512, 192, 1, bias=False)]),
            mask_constraint="linear", drop_first_bin=True, verbose=verbose)
    if name in ("cfg5", "cfg5_offline"):
        return SoTaskWrapModule(
            FreeEncDec(320, 512, 160),
            ConvTasNet(512, 0, False, tcn_dim=512, per_tcn_stack=8, repeat_tcn=3, tcn_with_embed=[0] * 8, tcn_norm="cLN",
                       dconv_norm="cLN", causal=True),
            mask_constraint="ReLU", verbose=verbose)
    if name in ("veve_dprnn_v0_causal", "tse_skim_v0", "tse_skim_v0_causal", "tse_skim_v1_causal", "tse_skim_v2_causal", "tse_skim_v0_causal_vad", "tse_unet_tcn_v0", "tse_unet_tcn_v0_causal", "tse_unet_tcn_v1",
                "ns_dpcrn_v0", "ns_dpcrn_v0_causal", "ns_dparn_v0", "ns_dparn_v0_causal"):
        return init_model(name, None, None, verbose=verbose)
    raise NameError(name)
