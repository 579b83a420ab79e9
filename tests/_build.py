"""Build this repo's drop-in modules from the plain-dict configs stored in the golden files."""
import torch.nn as nn

from puresound_b200.nnet.base_nn import SoTaskWrapModule
from puresound_b200.nnet.conv_tasnet import TCN, ConvTasNet, GatedTCN
from puresound_b200.nnet.dparn import DPARN
from puresound_b200.nnet.dpcrn import DPCRN
from puresound_b200.nnet.dprnn import DPRNN
from puresound_b200.nnet.lobe.encoder import ConvEncDec, FbankEnc, FreeEncDec
from puresound_b200.nnet.lobe.pooling import AttentiveStatisticsPooling
from puresound_b200.nnet.lobe.rnn import SingleRNN
from puresound_b200.nnet.lobe.trivial import Magnitude, SpecAugment
from puresound_b200.nnet.skim import SkiM
from puresound_b200.nnet.unet import UnetTcn


def tcn(c):
    return TCN(c["in_channels"], c["hid_channels"], c["kernel"], c["dilation"], emb_dim=c["emb_dim"], causal=c["causal"],
               tcn_norm=c["tcn_norm"], dconv_norm=c["dconv_norm"])


def gated_tcn(c):
    return GatedTCN(c["in_channels"], c["hid_channels"], c["kernel"], c["dilation"], emb_dim=c["emb_dim"], causal=c["causal"],
                    tcn_norm=c["tcn_norm"], use_film=c["use_film"])


def encoder(c):
    if c["type"] == "FreeEncDec":
        return FreeEncDec(c["win_length"], c["laten_length"], c["hop_length"], c["output_active"])
    if c["type"] == "FbankEnc":
        return FbankEnc(c["fft_length"], "hann", c["fft_length"], hop_length=c["hop_length"], trainable=c["trainable"],
                        output_format="Magnitude", n_banks=c["n_banks"])
    return ConvEncDec(c["fft_length"], "hann", c["win_length"], hop_length=c["hop_length"], trainable=True, output_format="Complex")


def masker(c):
    c = dict(c)
    t = c.pop("type")
    if t == "ConvTasNet":
        return ConvTasNet(**c)
    if t == "UnetTcn":
        return UnetTcn(**c)
    if t == "DPCRN":
        return DPCRN(**c)
    if t == "DPARN":
        return DPARN(**c)
    out = c.pop("output_size", c["input_size"])
    if t == "SkiM":
        return SkiM(c["input_size"], c["hidden_size"], out, n_blocks=c["n_blocks"], seg_size=c["seg_size"], seg_overlap=c["seg_overlap"],
                    causal=c["causal"], embed_dim=c["embed_dim"], embed_norm=c["embed_norm"], embed_fusion=c["embed_fusion"],
                    block_with_embed=c["block_with_embed"])
    return DPRNN(c["input_size"], c["hidden_size"], out, n_blocks=c["n_blocks"], seg_size=c["seg_size"], seg_overlap=c["seg_overlap"],
                 causal=c["causal"], embed_dim=c["embed_dim"], embed_norm=c["embed_norm"], block_with_embed=c["block_with_embed"],
                 embedding_free_tse=c["embedding_free_tse"])


def speaker_net(layers):
    mods = []
    for l in layers:
        t = l["type"]
        if t == "Magnitude":
            mods.append(Magnitude(l["drop_first"], l["log1p"]))
        elif t == "TCN":
            mods.append(tcn(l))
        elif t == "GatedTCN":
            mods.append(gated_tcn(l))
        elif t == "AttentiveStatisticsPooling":
            mods.append(AttentiveStatisticsPooling(l["channels"], l["attention_channels"]))
        elif t == "SpecAugment":
            mods.append(SpecAugment(l["freq_mask"], l["time_mask"], l["mask_value"]))
        elif t == "SingleRNN":
            mods.append("SingleRNN" if not l["bidirectional"] else "SingleRNN_bi")  # sized from the state dict by the caller
        elif t == "Conv1d":
            mods.append(None)  # sized from the state dict by the caller
    return mods


def wrapper(cfg, sd):
    spk = None
    if cfg["speaker_net"] is not None:
        mods = speaker_net(cfg["speaker_net"])
        for j, m in enumerate(mods):
            if m is None:
                w = sd[f"speaker_net.{j}.weight"]
                mods[j] = nn.Conv1d(w.shape[1], w.shape[0], 1, bias=f"speaker_net.{j}.bias" in sd)
            elif isinstance(m, str):
                w = sd[f"speaker_net.{j}.rnn.weight_hh_l0"]  # [4H, H]
                mods[j] = SingleRNN("LSTM", sd[f"speaker_net.{j}.rnn.weight_ih_l0"].shape[1], w.shape[1], bidirectional=m.endswith("_bi"))
        spk = nn.ModuleList(mods)
    m = SoTaskWrapModule(
        encoder(cfg["encoder"]), masker(cfg["masker"]), embedding_free_tse=cfg["embedding_free_tse"],
        encoder_spk=None if cfg["encoder_spk"] is None else encoder(cfg["encoder_spk"]), speaker_net=spk,
        f_type=cfg["f_type"], mask_type=cfg["mask_type"], mask_constraint=cfg["mask_constraint"],
        output_constraint=cfg["output_constraint"], drop_first_bin=cfg["drop_first_bin"], verbose=False)
    m.load_state_dict(sd, strict=True)
    return m.eval()
