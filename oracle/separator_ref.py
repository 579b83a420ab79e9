"""CPU oracle: a functional restatement of PureSound's separator forward pass.

TEST INFRASTRUCTURE — see ``oracle/__init__.py`` for who may import this.

The reference (mcw519/PureSound) is pure Python on top of PyTorch, so the
arithmetic of the path lives in a third-party dependency, ``torch`` (ATen /
oneDNN; ``requirements.txt:3`` unpinned, the container has 2.11.0+cu128).  This
file restates the path as *stateless functions over a reference-format
state_dict* (same keys as SURVEY.md appendix B), calling the same
``torch.nn.functional`` primitives at the same call sites the reference does,
in fp32 on the CPU.  Every function cites the reference file:line it follows
(paths relative to the reference root).

Parity pinning: the reference holds NO numerical golden vectors for this path
(only shape tests, plus SplitMerge round-trip and streaming==offline pins).  The
oracle is therefore pinned against *outputs of the reference itself*:
``tests/golden/make_golden.py`` imports the reference in the authoring
container, runs it on seeded weights/inputs and commits the tensors as
``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` replays them through
this file.  ``tests/golden/full_size_pins.json`` additionally holds reference
outputs of the full-size BASELINE configs (sub-sampled) on seeded weights.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------- #
# A1  learned filterbank                       puresound/nnet/lobe/encoder.py
# --------------------------------------------------------------------------- #
def free_encode(wav: Tensor, weight: Tensor, hop: int, relu: bool = False) -> Tensor:
    """FreeEncDec.forward, lobe/encoder.py:71-83 (conv built at :50-56).

    wav [N, L] -> feats [N, Nf, T], T = floor((L - win) / hop) + 1, no bias,
    no padding, optional ReLU (``output_active``)."""
    y = F.conv1d(wav.unsqueeze(1), weight, stride=hop)
    return F.relu(y) if relu else y


def free_decode(feats: Tensor, weight: Tensor, hop: int) -> Tensor:
    """FreeEncDec.inverse, lobe/encoder.py:85-94 (ConvTranspose1d at :62-68).

    feats [N, Nf, T] -> wav [N, (T-1)*hop + win]: plain overlap-add sum."""
    return F.conv_transpose1d(feats, weight, stride=hop).squeeze(1)


# --------------------------------------------------------------------------- #
# A2/A3  conv-STFT and conv-iSTFT   lobe/encoder.py:275-456, lobe/stft.py:8-125
# --------------------------------------------------------------------------- #
def fourier_kernels(n_fft: int) -> Tuple[Tensor, Tensor]:
    """create_fourier_kernels(freq_scale='no'), lobe/stft.py:91-100: unwindowed
    sin/cos kernels [n_fft//2+1, 1, n_fft], built in float64 then cast."""
    k = torch.arange(n_fft // 2 + 1, dtype=torch.float64).view(-1, 1)
    s = torch.arange(n_fft, dtype=torch.float64).view(1, -1)
    ang = 2.0 * math.pi * k * s / n_fft
    return (torch.sin(ang).float().unsqueeze(1), torch.cos(ang).float().unsqueeze(1))


def stft_encode(wav: Tensor, wsin: Tensor, wcos: Tensor, hop: int) -> Tensor:
    """ConvSTFT.forward (output_format='Complex'), lobe/encoder.py:358-378.

    wav [N, L] -> [N, F, T, 2] with last axis (real, -imag_conv)."""
    x = wav.unsqueeze(1)
    im = F.conv1d(x, wsin, stride=hop)
    re = F.conv1d(x, wcos, stride=hop)
    return torch.stack((re, -im), dim=-1)


def mel_encode(wav: Tensor, wsin: Tensor, wcos: Tensor, filterbank: Tensor, hop: int, trainable: bool) -> Tensor:
    """FbankEnc.forward -> ConvMelSpectrogram.forward with output_format='Magnitude', lobe/encoder.py:249-259,509-535:
    POWER spectrum (no square root; + 1e-8 when trainable) times the mel filterbank [bins, n_banks].  [N, L] -> [N, n_banks, T]."""
    x = wav.unsqueeze(1)
    im = F.conv1d(x, wsin, stride=hop)
    re = F.conv1d(x, wcos, stride=hop)
    spec = re.pow(2) + im.pow(2)
    if trainable:
        spec = spec + 1e-8
    return torch.matmul(spec.permute(0, 2, 1), filterbank).permute(0, 2, 1)


def stft_decode(
    X: Tensor, kernel_cos_inv: Tensor, kernel_sin_inv: Tensor, window_mask: Tensor, hop: int, n_fft: int
) -> Tensor:
    """ConvSTFT.inverse, lobe/encoder.py:393-456 with extend_fbins
    (lobe/stft.py:118-125), overlap_add (:103-106) and torch_window_sumsquare
    (:109-115).  X [N, F, T, 2] -> wav [N, n_fft + hop*(T-1)]."""
    upper = X[:, 1:-1].flip(1).clone()
    upper[..., 1] = -upper[..., 1]
    Xf = torch.cat((X, upper), dim=1)  # [N, n_fft, T, 2]
    a1 = F.conv2d(Xf[..., 0].unsqueeze(1), kernel_cos_inv, stride=(1, 1))
    b2 = F.conv2d(Xf[..., 1].unsqueeze(1), kernel_sin_inv, stride=(1, 1))
    real = (a1 - b2).squeeze(-2) * window_mask
    real = real / n_fft
    T = X.shape[2]
    out_len = n_fft + hop * (T - 1)
    real = F.fold(real, (1, out_len), kernel_size=(1, n_fft), stride=hop).flatten(1)
    w = window_mask.flatten()
    w_stack = (w.unsqueeze(-1).repeat(1, T) ** 2).unsqueeze(0)
    w_sum = F.fold(w_stack, (1, out_len), kernel_size=(1, n_fft), stride=hop).flatten()
    nz = w_sum > 1e-10
    real[:, nz] = real[:, nz].div(w_sum[nz])
    return real


# --------------------------------------------------------------------------- #
# A4  norms                                       puresound/nnet/lobe/norm.py
# --------------------------------------------------------------------------- #
def _gain_bias(x: Tensor, g: Tensor, b: Tensor) -> Tensor:
    """_LayerNorm.apply_gain_and_bias, lobe/norm.py:15-17."""
    return (g * x.transpose(1, -1) + b).transpose(1, -1)


def norm_apply(kind: str, sd: SD, p: str, x: Tensor) -> Tensor:
    """get_norm registry, lobe/norm.py:100-112.  x [N, C, T]."""
    if kind == "gLN":  # GlobLN.forward, lobe/norm.py:23-34
        dims = list(range(1, x.dim()))
        mean = x.mean(dim=dims, keepdim=True)
        var = torch.pow(x - mean, 2).mean(dim=dims, keepdim=True)
        return _gain_bias((x - mean) / (var + 1e-8).sqrt(), sd[p + "gamma"], sd[p + "beta"])
    if kind == "cLN":  # ChanLN.forward, lobe/norm.py:40-50
        mean = torch.mean(x, dim=1, keepdim=True)
        var = torch.var(x, dim=1, keepdim=True, unbiased=False)
        return _gain_bias((x - mean) / (var + 1e-8).sqrt(), sd[p + "gamma"], sd[p + "beta"])
    if kind == "gGN":  # lobe/norm.py:96  GroupNorm(1, C, 1e-8)
        return F.group_norm(x, 1, sd[p + "weight"], sd[p + "bias"], 1e-8)
    if kind == "bN1d":  # lobe/norm.py:94, eval mode (running statistics)
        return F.batch_norm(
            x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False, 0.0, 1e-5
        )
    raise NameError("Could not interpret normalization identifier")


# --------------------------------------------------------------------------- #
# A5/A6  TCN block and Conv-TasNet stack     puresound/nnet/conv_tasnet.py
# --------------------------------------------------------------------------- #
def tcn_block(
    sd: SD,
    p: str,
    x: Tensor,
    embed: Optional[Tensor],
    kernel: int,
    dilation: int,
    causal: bool,
    tcn_norm: str,
    dconv_norm: str,
) -> Tensor:
    """TCN.forward, conv_tasnet.py:67-90, with DepthwiseSeparableConv1d.forward
    (lobe/cnn.py:84-106; padding rule :58-60, causal trim :100-101)."""
    res = x
    if embed is not None:
        e = embed.unsqueeze(2).repeat(1, 1, x.size(2))
        x = torch.cat([x, e], dim=1)
    x = F.conv1d(x, sd[p + "in_conv.0.weight"])
    x = norm_apply(tcn_norm, sd, p + "in_conv.1.", x)
    x = F.prelu(x, sd[p + "in_conv.2.weight"])
    d = p + "dconv.0."
    pad = (kernel - 1) * dilation if causal else ((kernel - 1) // 2) * dilation
    x = F.conv1d(
        x, sd[d + "depthwise.0.weight"], sd[d + "depthwise.0.bias"], dilation=dilation, padding=pad, groups=x.size(1)
    )
    x = norm_apply(dconv_norm, sd, d + "depthwise.1.", x)
    x = F.prelu(x, sd[d + "depthwise.2.weight"])
    x = F.conv1d(x, sd[d + "pointwise.0.weight"], sd[d + "pointwise.0.bias"])
    x = norm_apply(dconv_norm, sd, d + "pointwise.1.", x)
    x = F.prelu(x, sd[d + "pointwise.2.weight"])
    if causal:
        x = x[..., :-pad]
    x = F.conv1d(x, sd[p + "out_conv.weight"], sd[p + "out_conv.bias"])
    return x + res


def gated_tcn_block(
    sd: SD, p: str, x: Tensor, embed: Optional[Tensor], kernel: int, dilation: int, causal: bool, tcn_norm: str
) -> Tensor:
    """GatedTCN.forward, conv_tasnet.py:174-215.  Padding rule :122-124: both sides by ``padd`` (the
    causal variant pads (k-1)*d on both sides and trims the tail after out_conv, :209-210); the embedding is concatenated
    before the zero padding of the right conv (:191-194) or applied as FiLM (:196-200, present iff the block holds
    ``cond_scale``)."""
    pad = (kernel - 1) * dilation if causal else (kernel - 1) * dilation // 2
    res = x
    x = F.conv1d(x, sd[p + "in_conv.weight"])
    if embed is not None:
        if p + "cond_scale.weight" not in sd:
            x_r = torch.cat([x, embed.unsqueeze(-1).repeat(1, 1, x.size(2))], dim=1)
        else:
            c = embed.unsqueeze(-1)
            x_r = F.conv1d(c, sd[p + "cond_scale.weight"]) * x + F.conv1d(c, sd[p + "cond_bias.weight"])
    else:
        x_r = x
    left = F.conv1d(x, sd[p + "left_conv.0.weight"], dilation=dilation, padding=pad)
    left = F.prelu(norm_apply(tcn_norm, sd, p + "left_conv.1.", left), sd[p + "left_conv.2.weight"])
    right = F.conv1d(x_r, sd[p + "right_conv.0.weight"], dilation=dilation, padding=pad)
    right = torch.sigmoid(F.prelu(norm_apply(tcn_norm, sd, p + "right_conv.1.", right), sd[p + "right_conv.2.weight"]))
    x = F.conv1d(left * right, sd[p + "out_conv.weight"])
    if causal:
        x = x[..., :-pad]
    return x + res


def conv_tasnet(sd: SD, p: str, x: Tensor, dvec: Optional[Tensor], a: dict) -> Tensor:
    """ConvTasNet.forward, conv_tasnet.py:338-359; dilation schedule ``tcn_dilated_basic ** i`` from :290; block class
    by ``tcn_layer`` (:270-275; GatedTCN takes no dconv_norm)."""
    layer = a["tcn_layer"].lower()
    if layer not in ("normal", "gated"):
        raise NameError
    if a["embed_norm"] and dvec is not None:
        dvec = F.normalize(dvec, p=2, dim=1)
    for r in range(a["repeat_tcn"]):
        for i in range(a["per_tcn_stack"]):
            q, e, d = f"{p}tcn_list.{r}.{i}.", dvec if a["tcn_with_embed"][i] else None, a["tcn_dilated_basic"] ** i
            if layer == "gated":
                x = gated_tcn_block(sd, q, x, e, a["tcn_kernel"], d, a["causal"], a["tcn_norm"])
            else:
                x = tcn_block(sd, q, x, e, a["tcn_kernel"], d, a["causal"], a["tcn_norm"], a["dconv_norm"])
    return x


# --------------------------------------------------------------------------- #
# A7  segmentation                     lobe/trivial.py:170-241, dprnn.py:133-145
# --------------------------------------------------------------------------- #
def split_overlap(x: Tensor, K: int) -> Tuple[Tensor, int]:
    """SplitMerge.split, lobe/trivial.py:178-210.  [N,C,T] -> ([N,S,K,C], rest)."""
    s = K // 2
    N, C, T = x.shape
    rest = K - (s + T % K) % K
    if rest > 0:
        x = torch.cat([x, x.new_zeros(N, C, rest)], dim=-1)
    z = x.new_zeros(N, C, s)
    x = torch.cat([z, x, z], dim=-1)
    a = x[:, :, :-s].contiguous().view(N, C, -1, K)
    b = x[:, :, s:].contiguous().view(N, C, -1, K)
    seg = torch.cat([a, b], dim=-1).view(N, C, -1, K)
    return seg.permute(0, 2, 3, 1), rest


def merge_overlap(x: Tensor, rest: int) -> Tensor:
    """SplitMerge.merge, lobe/trivial.py:212-241.  [N,S,K,C] -> [N,C,T]."""
    N, S, K, C = x.shape
    s = K // 2
    x = x.permute(0, 3, 1, 2).contiguous().view(N, C, -1, 2 * K)
    x1 = x[:, :, :, :K].contiguous().view(N, C, -1)[:, :, s:]
    x2 = x[:, :, :, K:].contiguous().view(N, C, -1)[:, :, :-s]
    out = (x1 + x2) / 2
    if rest > 0:
        out = out[..., :-rest]
    return out.contiguous()


# --------------------------------------------------------------------------- #
# A8/A9  LSTM, FiLM, DPRNN                         puresound/nnet/dprnn.py
# --------------------------------------------------------------------------- #
def _lstm_dir(x: Tensor, w_ih, w_hh, b_ih, b_hh, h, c, reverse: bool):
    """One direction of torch.nn.LSTM (gate order i,f,g,o) written out step by
    step: the published cell the reference reaches through nn.LSTM
    (dprnn.py:67-103)."""
    B, L, _ = x.shape
    gx = F.linear(x, w_ih, b_ih)
    out = x.new_empty(B, L, w_hh.shape[1])
    steps = range(L - 1, -1, -1) if reverse else range(L)
    for t in steps:
        g = gx[:, t] + F.linear(h, w_hh, b_hh)
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[:, t] = h
    return out, h, c


def lstm(
    sd: SD, p: str, x: Tensor, bidirectional: bool, init: Optional[Tuple[Tensor, Tensor]] = None, fast: bool = True
) -> Tuple[Tensor, Tuple[Tensor, Tensor]]:
    """nn.LSTM(num_layers=1, batch_first=True) over x [B, L, C] with reference
    keys ``weight_ih_l0[_reverse]`` etc.  ``fast`` routes through ATen's fused
    LSTM (what the reference itself executes); ``fast=False`` is the explicit
    cell loop above.  Tests check the two agree."""
    D = 2 if bidirectional else 1
    H = sd[p + "weight_hh_l0"].shape[1]
    B = x.shape[0]
    if init is None:
        h0 = x.new_zeros(D, B, H)
        c0 = x.new_zeros(D, B, H)
    else:
        h0, c0 = init
    sfx = ["", "_reverse"][:D]
    if fast:
        flat = []
        for s in sfx:
            flat += [sd[p + f"weight_ih_l0{s}"], sd[p + f"weight_hh_l0{s}"], sd[p + f"bias_ih_l0{s}"], sd[p + f"bias_hh_l0{s}"]]
        out, hn, cn = torch._VF.lstm(x, (h0, c0), flat, True, 1, 0.0, False, bidirectional, True)
        return out, (hn, cn)
    outs, hs, cs = [], [], []
    for d, s in enumerate(sfx):
        o, h, c = _lstm_dir(
            x,
            sd[p + f"weight_ih_l0{s}"],
            sd[p + f"weight_hh_l0{s}"],
            sd[p + f"bias_ih_l0{s}"],
            sd[p + f"bias_hh_l0{s}"],
            h0[d],
            c0[d],
            reverse=(d == 1),
        )
        outs.append(o)
        hs.append(h)
        cs.append(c)
    return torch.cat(outs, dim=-1), (torch.stack(hs), torch.stack(cs))


def film(sd: SD, p: str, x: Tensor, cond: Tensor, input_norm: bool = True) -> Tensor:
    """FiLM.forward, lobe/trivial.py:148-167.  x [N', C, K], cond [N', E]."""
    if input_norm:
        C = x.shape[1]
        x = F.layer_norm(x.transpose(1, 2), (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5).transpose(1, 2)
    q = torch.cat([x, cond.unsqueeze(-1).repeat(1, 1, x.shape[-1])], dim=1)
    return F.conv1d(q, sd[p + "cond_scale.weight"]) * x + F.conv1d(q, sd[p + "cond_bias.weight"])


def _dprnn_segment(x: Tensor, K: int, overlap: bool) -> Tuple[Tensor, int]:
    """dprnn.py:133-145 (and the identical :205-217)."""
    if overlap:
        return split_overlap(x, K)
    N, C, T = x.shape
    x = x.permute(0, 2, 1)
    rest = K - T % K
    if rest > 0:
        x = F.pad(x, (0, 0, 0, rest))
    return x.reshape(N, -1, K, C), rest


def _dprnn_blocks(sd: SD, p: str, seg: Tensor, a: dict, embed, inits, collect_hidden: bool, fast_lstm: bool):
    """The block loop shared by DPRNN.forward (dprnn.py:153-178) and
    DPRNN._get_hidden_states (:221-244)."""
    N, S, K, C = seg.shape
    bi = not a["causal"]
    out = seg
    hidden = []
    for i in range(a["n_blocks"]):
        out = out.reshape(-1, K, C).contiguous()
        if embed is not None and a["block_with_embed"] is not None and a["block_with_embed"][i]:
            out = film(sd, f"{p}input_film.{i}.", out.transpose(1, 2), embed).transpose(1, 2)
        y, _ = lstm(sd, f"{p}intra_rnn.{i}.", out, bi, None, fast_lstm)
        y = F.linear(y, sd[f"{p}intra_proj.{i}.weight"], sd[f"{p}intra_proj.{i}.bias"])
        y = F.layer_norm(y, (C,), sd[f"{p}intra_norm.{i}.weight"], sd[f"{p}intra_norm.{i}.bias"], 1e-5)
        out = out + y
        v = out.reshape(N, S, K, C).permute(0, 2, 1, 3).reshape(-1, S, C).contiguous()
        y, hid = lstm(sd, f"{p}inter_rnn.{i}.", v, bi, inits[i], fast_lstm)
        hidden.append(hid)
        y = F.linear(y, sd[f"{p}inter_proj.{i}.weight"], sd[f"{p}inter_proj.{i}.bias"])
        y = F.layer_norm(y, (C,), sd[f"{p}inter_norm.{i}.weight"], sd[f"{p}inter_norm.{i}.bias"], 1e-5)
        out = v + y
        out = out.reshape(N, K, S, C).contiguous().permute(0, 2, 1, 3)
    return (hidden if collect_hidden else out)


def dprnn_hidden_states(sd: SD, p: str, x: Tensor, a: dict, fast_lstm: bool = True):
    """DPRNN._get_hidden_states, dprnn.py:193-244."""
    seg, _ = _dprnn_segment(x, a["seg_size"], a["seg_overlap"])
    return _dprnn_blocks(sd, p, seg, a, None, [None] * a["n_blocks"], True, fast_lstm)


def dprnn(sd: SD, p: str, x: Tensor, embed: Optional[Tensor], a: dict, fast_lstm: bool = True) -> Tensor:
    """DPRNN.forward, dprnn.py:111-191."""
    if a["embedding_free_tse"]:
        assert embed is not None and embed.dim() == 3, "embedding free tse need enrollment waveform as input."
        inits = dprnn_hidden_states(sd, p, embed, a, fast_lstm)
    else:
        inits = [None] * a["n_blocks"]
    if a["embed_norm"] and embed is not None and not a["embedding_free_tse"]:
        embed = F.normalize(embed, p=2, dim=1)
    N, C, T = x.shape
    seg, rest = _dprnn_segment(x, a["seg_size"], a["seg_overlap"])
    S = seg.shape[1]
    film_embed = None
    if not a["embedding_free_tse"] and embed is not None:
        film_embed = embed.unsqueeze(1).repeat(1, S, 1).reshape(N * S, -1)
    out = _dprnn_blocks(sd, p, seg, a, film_embed, inits, False, fast_lstm)
    if a["seg_overlap"]:
        out = merge_overlap(out.reshape(N, S, a["seg_size"], C), rest)
    else:
        out = out.reshape(N, S * a["seg_size"], C)[:, :T, :].transpose(1, 2)
    out = F.prelu(out, sd[p + "output_fc.0.weight"])
    return F.conv1d(out, sd[p + "output_fc.1.weight"], sd[p + "output_fc.1.bias"])


# --------------------------------------------------------------------------- #
# F2  SkiM (skipping-memory LSTM)                       puresound/nnet/skim.py
# --------------------------------------------------------------------------- #
def _skim_split(x: Tensor, K: int) -> Tuple[Tensor, int]:
    """SkiM.split, skim.py:349-383.  [N,C,T] -> ([N,S,K,C], rest): the same construction as SplitMerge.split."""
    return split_overlap(x, K)


def _seg_lstm(sd: SD, p: str, x: Tensor, h, c, bi: bool, fast_lstm: bool):
    """SegLSTM.forward, skim.py:198-229: x [NS,K,C], (h,c) [D,NS,H] or None -> (x + LN(proj(LSTM(x)))), h_n, c_n."""
    H = sd[p + "lstm.weight_hh_l0"].shape[1]
    D = 2 if bi else 1
    if h is None:
        h = x.new_zeros(D, x.shape[0], H)
    if c is None:
        c = x.new_zeros(D, x.shape[0], H)
    y, (hn, cn) = lstm(sd, p + "lstm.", x, bi, (h, c), fast_lstm)
    y = F.linear(y, sd[p + "proj.weight"], sd[p + "proj.bias"])
    y = F.layer_norm(y, (x.shape[2],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return x + y, hn, cn


def _mem_lstm(sd: SD, p: str, h: Tensor, c: Tensor, causal: bool, fast_lstm: bool):
    """MemLSTM.forward (streaming=False), skim.py:46-117: h, c [N,S,D,H] -> ([D,NS,H], [D,NS,H]) for the next SegLSTM;
    in the causal setting segment s receives the memory of segment s-1 (zeros for the first)."""
    N, S, D, H = h.shape
    out = []
    for name, v in (("h", h), ("c", c)):
        v = v.reshape(N, S, D * H)
        y, _ = lstm(sd, f"{p}{name}_net.", v, not causal, None, fast_lstm)
        y = F.linear(y.reshape(N * S, -1), sd[f"{p}{name}_proj.weight"], sd[f"{p}{name}_proj.bias"]).reshape(N, S, -1)
        v = v + F.layer_norm(y, (D * H,), sd[f"{p}{name}_norm.weight"], sd[f"{p}{name}_norm.bias"], 1e-5)
        v = v.reshape(N * S, D, H).transpose(1, 0).contiguous()  # [D, NS, H]
        if causal:
            z = torch.zeros_like(v)
            z[:, 1:, :] = v[:, :-1, :]
            v = z
        out.append(v)
    return out[0], out[1]


def skim(sd: SD, p: str, x: Tensor, embed: Optional[Tensor], a: dict, fast_lstm: bool = True) -> Tensor:
    """SkiM.forward, skim.py:414-469.  x [N,C,T], embed [N,E] -> [N,C_out,T]."""
    if a["embed_norm"] and embed is not None:
        embed = F.normalize(embed, p=2, dim=1)
    K, bi = a["seg_size"], not a["causal"]
    N, C, T = x.shape
    if a["seg_overlap"]:
        seg, rest = _skim_split(x, K)
    else:
        xt = x.permute(0, 2, 1)
        rest = K - T % K
        if rest > 0:
            xt = F.pad(xt, (0, 0, 0, rest))
        seg = xt.reshape(N, -1, K, C)
    S = seg.shape[1]
    e = None
    if embed is not None:
        e = embed.unsqueeze(1).repeat(1, S, 1).reshape(N * S, -1)
    out = seg.reshape(N * S, K, C).contiguous()
    h = c = None
    H = a["hidden_size"]
    for i in range(a["n_blocks"]):
        if e is not None and a["block_with_embed"][i]:
            out = film(sd, f"{p}seg_input_fusion.{i}.", out.transpose(1, 2), e).transpose(1, 2)
        out, h, c = _seg_lstm(sd, f"{p}seg_lstm.{i}.", out, h, c, bi, fast_lstm)
        if i < a["n_blocks"] - 1:
            h = h.reshape(-1, N, S, H).permute(1, 2, 0, 3)
            c = c.reshape(-1, N, S, H).permute(1, 2, 0, 3)
            h, c = _mem_lstm(sd, f"{p}mem_lstm.{i}.", h, c, a["causal"], fast_lstm)
    if a["seg_overlap"]:
        out = merge_overlap(out.reshape(N, S, K, C), rest)
    else:
        out = out.reshape(N, S * K, C)[:, :T, :].transpose(1, 2)
    out = F.prelu(out, sd[p + "output_fc.0.weight"])
    return F.conv1d(out, sd[p + "output_fc.1.weight"], sd[p + "output_fc.1.bias"])


# --------------------------------------------------------------------------- #
# A10  speaker path            lobe/trivial.py:21-58, lobe/pooling.py:58-126
# --------------------------------------------------------------------------- #
def magnitude(x: Tensor, drop_first: bool = True, log1p: bool = False) -> Tensor:
    """Magnitude.forward on the 3-D channel-cat layout, lobe/trivial.py:35-58."""
    re, im = torch.chunk(x, 2, dim=1)
    if drop_first:
        re, im = re[:, 1:, :], im[:, 1:, :]
    mag = torch.sqrt(re.pow(2) + im.pow(2) + 1e-8)
    return torch.log1p(mag) if log1p else mag


def asp(sd: SD, p: str, x: Tensor) -> Tensor:
    """AttentiveStatisticsPooling.forward with lengths=None (all frames valid,
    which is how every caller uses it), eval-mode BatchNorm,
    lobe/pooling.py:87-126.  x [N, C, T] -> [N, 2C, 1]."""
    a = F.conv1d(x, sd[p + "tdnn.0.weight"], sd[p + "tdnn.0.bias"])
    a = F.relu(a)
    a = F.batch_norm(
        a, sd[p + "tdnn.2.running_mean"], sd[p + "tdnn.2.running_var"], sd[p + "tdnn.2.weight"], sd[p + "tdnn.2.bias"], False, 0.0, 1e-5
    )
    a = F.conv1d(torch.tanh(a), sd[p + "conv.weight"], sd[p + "conv.bias"])
    w = F.softmax(a, dim=2)
    mean = (w * x).sum(2)
    std = torch.sqrt((w * (x - mean.unsqueeze(2)).pow(2)).sum(2).clamp(1e-12))
    return torch.cat((mean, std), dim=1).unsqueeze(2)


def spec_augment(x: Tensor, freq_mask: int, time_mask: int, value: float) -> Tensor:
    """SpecAugment.forward, lobe/trivial.py:324-335 - applied in eval mode too (there is no `self.training` test upstream).
    Restates torchaudio.functional.mask_along_axis (torchaudio 2.11, the function trivial.py:7 imports): per axis with a
    mask length >= 1, `value = rand(1) * mask_param`, `min_value = rand(1) * (axis_len - value)` from the GLOBAL CPU generator,
    band [long(min_value), long(min_value) + long(value)) filled for every item alike.  x [N, C, T]."""
    for axis, param in ((1, freq_mask), (2, time_mask)):
        if param < 1:
            continue
        v = torch.rand(1) * param
        v0 = torch.rand(1) * (x.shape[axis] - v)
        start, end = int(v0.long()), int(v0.long()) + int(v.long())
        idx = torch.arange(x.shape[axis])
        band = (idx >= start) & (idx < end)
        x = x.masked_fill(band.view(1, -1, 1) if axis == 1 else band.view(1, 1, -1), value)
    return x


def single_rnn(sd: SD, p: str, x: Tensor, bidirectional: bool, fast_lstm: bool = True) -> Tensor:
    """SingleRNN.forward (LSTM), lobe/rnn.py:36-52: x [N, C, T] -> LSTM over T -> dropout (eval) -> Linear -> [N, C, T]."""
    h, _ = lstm(sd, p + "rnn.", x.permute(0, 2, 1), bidirectional, fast=fast_lstm)
    return F.linear(h, sd[p + "proj.weight"], sd[p + "proj.bias"]).permute(0, 2, 1)


def speaker_net(sd: SD, p: str, layers: Sequence[dict], x: Tensor) -> Tensor:
    """The ModuleList speaker nets of the TSE recipes (egs/tse/model.py:118-135),
    applied layer by layer as base_nn.py:699-705 does.  Returns [N, E]."""
    for j, l in enumerate(layers):
        q = f"{p}{j}."
        t = l["type"]
        if t == "Magnitude":
            x = magnitude(x, l["drop_first"], l["log1p"])
        elif t == "TCN":
            x = tcn_block(sd, q, x, None, l["kernel"], l["dilation"], l["causal"], l["tcn_norm"], l["dconv_norm"])
        elif t == "GatedTCN":  # the speaker nets of the tse_unet_tcn recipes (egs/tse/model.py:226-236)
            x = gated_tcn_block(sd, q, x, None, l["kernel"], l["dilation"], l["causal"], l["tcn_norm"])
        elif t == "AttentiveStatisticsPooling":
            x = asp(sd, q, x)
        elif t == "SpecAugment":  # speaker net of tse_skim_v2_causal (egs/tse/model.py:536)
            x = spec_augment(x, l["freq_mask"], l["time_mask"], l["mask_value"])
        elif t == "SingleRNN":  # speaker net of tse_skim_v1_causal (egs/tse/model.py:489-499)
            x = single_rnn(sd, q, x, l["bidirectional"])
        elif t == "Conv1d":
            x = F.conv1d(x, sd[q + "weight"], sd.get(q + "bias"))
        else:
            raise NotImplementedError(t)
    return x.squeeze(-1)


# --------------------------------------------------------------------------- #
# A6/A11/A12  task wrapper                       puresound/nnet/base_nn.py
# --------------------------------------------------------------------------- #
def get_mask(mask: Tensor, constraint: str) -> Tensor:
    """EncDecMaskerBaseModel.get_mask, base_nn.py:81-95."""
    c = constraint.lower()
    if c == "linear":
        return mask
    if c == "relu":
        return torch.relu(mask)
    if c == "sigmoid":
        return torch.sigmoid(mask)
    raise NotImplementedError


def apply_tf_masks(tf: Tensor, m: Tensor, mask_type: str, f_type: str) -> Tensor:
    """apply_tf_masks, base_nn.py:41-79 — the two working combinations
    (real/real :146-159 and complex/complex :97-112); output re-laid as the
    channel-cat [N, 2F, T] that _get_waveform re-stacks (:381-383)."""
    mt, ft = mask_type.lower(), f_type.lower()
    if mt == "real" and ft == "real":
        return tf * m
    if mt == "complex" and ft == "complex":
        a, b = torch.chunk(tf, 2, dim=1)
        c, d = torch.chunk(m, 2, dim=1)
        return torch.cat([a * c - b * d, a * d + b * c], dim=1)
    if (mt, ft) in (("real", "complex"), ("polar", "polar")):
        raise NotImplementedError("broken upstream (base_nn.py:127, :75); not targeted")
    raise NameError


def _encode(sd: SD, p: str, e: dict, wav: Tensor, drop_first_bin: bool) -> Tensor:
    """SoTaskWrapModule._get_feature for one input, base_nn.py:335-345."""
    if e["type"] == "FreeEncDec":
        return free_encode(wav, sd[p + "encoder.weight"], e["hop_length"], e["output_active"])
    if e["type"] == "ConvEncDec":
        X = stft_encode(wav, sd[p + "encoder.wsin"], sd[p + "encoder.wcos"], e["hop_length"])
        re, im = X[..., 0], X[..., 1]
        if drop_first_bin:
            re, im = re[:, 1:, :], im[:, 1:, :]
        return torch.cat([re, im], dim=1)
    if e["type"] == "FbankEnc":
        return mel_encode(wav, sd[p + "encoder.wsin"], sd[p + "encoder.wcos"], sd[p + "encoder.filterbank"], e["hop_length"], e["trainable"])
    raise NotImplementedError(e["type"])


def _decode(sd: SD, p: str, e: dict, feats: Tensor, drop_first_bin: bool) -> Tensor:
    """SoTaskWrapModule._get_waveform, base_nn.py:379-396."""
    if e["type"] == "FreeEncDec":
        return free_decode(feats, sd[p + "decoder.weight"], e["hop_length"])
    re, im = torch.chunk(feats, 2, dim=1)
    if drop_first_bin:
        z = re.new_zeros(re.shape[0], 1, re.shape[2])
        re, im = torch.cat([z, re], dim=1), torch.cat([z, im], dim=1)
    X = torch.stack([re, im], dim=-1)
    return stft_decode(
        X, sd[p + "encoder.kernel_cos_inv"], sd[p + "encoder.kernel_sin_inv"], sd[p + "encoder.window_mask"], e["hop_length"], e["fft_length"]
    )


def wav_constrain(wav: Tensor, mode: str) -> Tensor:
    """_wav_output_constrain, base_nn.py:414-424."""
    m = mode.lower()
    if m == "linear":
        return torch.clamp(wav, min=-1, max=1)
    if m == "sigmoid":
        return torch.sigmoid(wav)
    raise NameError("Non support type.")


# --------------------------------------------------------------------------- #
# U-Net shell with a TCN bottleneck                 puresound/nnet/unet.py
# --------------------------------------------------------------------------- #
def _norm2d(kind: str, sd: SD, p: str, x: Tensor) -> Tensor:
    """The 2-D norms of the U-Net recipes on [N, C, F, T]: gLN (GlobLN over all non-batch dims, lobe/norm.py:20-34) and
    bN2d (nn.BatchNorm2d in eval mode, lobe/norm.py:95)."""
    if kind == "gLN":
        return norm_apply("gLN", sd, p, x)
    if kind == "bN2d":
        return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False, 0.0, 1e-5)
    raise NotImplementedError(kind)


def _unet_shell(sd: SD, p: str, x: Tensor, a: dict, bottleneck) -> Tensor:
    """The U-Net shell shared by Unet / UnetTcn / DPCRN (unet.py:219-273,454-517; dpcrn.py:136-190; layers built at
    unet.py:100-175): ZeroPad2d + Conv2d + norm + PReLU down path, ``bottleneck`` on [N, ch, F', T], cat-skip +
    ConvTranspose2d (+ norm + PReLU) up path with the time trim of unet.py:529-537."""
    n_cnn = len(a["kernel_t"])
    if a["input_type"].lower() == "ri":
        re, im = torch.chunk(x, 2, dim=-2)
        x = torch.stack([re, im], dim=1)
    else:
        x = x.unsqueeze(1)
    skip = [x]
    for i in range(n_cnn):
        kf, kt = a["kernel_f"][i], a["kernel_t"][i]
        q = f"{p}cnn_down.{i}."
        x = F.pad(x, (kt - a["delay"][i] - 1, a["delay"][i], kf // 2, kf // 2))
        x = F.conv2d(x, sd[q + "1.weight"], sd[q + "1.bias"], stride=(a["stride_f"][i], a["stride_t"][i]))
        x = F.prelu(_norm2d(a["norm_type"], sd, q + "2.", x), sd[q + "3.weight"])
        skip.append(x)
    x = bottleneck(x)
    tk = a["transpose_t_size"]
    for i in range(n_cnn):
        idx = n_cnn - 1 - i
        k, s = a["kernel_f"][idx], a["stride_f"][idx]
        q = f"{p}cnn_up.{i}."
        x = torch.cat([x, skip[-i - 1]], dim=1)
        x = F.conv_transpose2d(x, sd[q + "0.weight"], sd[q + "0.bias"], stride=(s, a["stride_t"][idx]), padding=(k // 2, 0),
                               output_padding=(s - k + 2 * (k // 2), 0))
        if idx != 0:
            x = F.prelu(_norm2d(a["norm_type"], sd, q + "1.", x), sd[q + "2.weight"])
        if tk != 1:
            x = x[..., (tk - 1):] if a["transpose_delay"] else x[..., : -(tk - 1)]
    if a["input_type"].lower() == "ri":
        return torch.cat([x[:, 0], x[:, 1]], dim=1)
    return x.squeeze(1)


def unet_tcn(sd: SD, p: str, x: Tensor, dvec: Optional[Tensor], a: dict) -> Tensor:
    """UnetTcn.forward, unet.py:454-517: the shell around a TCN / GatedTCN stack on the flattened [N, ch*F', T] tensor
    (stack built at unet.py:371-451)."""
    if a["embed_norm"] and dvec is not None:
        dvec = F.normalize(dvec, p=2, dim=1)

    def bottleneck(x):
        N, ch, Fb, T = x.shape
        x = x.reshape(N, ch * Fb, T)
        gated = a["tcn_layer"].lower() == "gated"
        for r in range(a["repeat_tcn"]):
            for i in range(a["per_tcn_stack"]):
                q, e, d = f"{p}tcn_list.{r}.{i}.", dvec if a["tcn_with_embed"][i] else None, a["tcn_dilated_basic"] ** i
                if gated:
                    x = gated_tcn_block(sd, q, x, e, a["tcn_kernel"], d, a["causal"], a["tcn_norm"])
                else:
                    x = tcn_block(sd, q, x, e, a["tcn_kernel"], d, a["causal"], a["tcn_norm"], a["dconv_norm"])
        return x.reshape(N, ch, Fb, T)

    return _unet_shell(sd, p, x, a, bottleneck)


def _dprnn_block2d(sd: SD, p: str, x: Tensor, fast_lstm: bool) -> Tensor:
    """DPRNNblock2D.forward, dpcrn.py:34-81 (SingleRNN = LSTM + Linear, lobe/rnn.py:36-52): bidirectional LSTM over the
    frequency rows of every frame, uni-directional LSTM over the frames of every frequency row, each followed by
    Linear -> LayerNorm(channels) -> + skip.  x [N, CH, C, T]."""
    N, CH, C, T = x.shape
    v = x.permute(0, 3, 2, 1).reshape(N * T, C, CH)
    y, _ = lstm(sd, p + "intra_rnn.rnn.", v, True, None, fast=fast_lstm)
    y = F.linear(y, sd[p + "intra_rnn.proj.weight"], sd[p + "intra_rnn.proj.bias"])
    y = F.layer_norm(y, (CH,), sd[p + "intra_norm.weight"], sd[p + "intra_norm.bias"], 1e-5)
    x = x + y.reshape(N, T, C, CH).permute(0, 3, 2, 1)
    v = x.permute(0, 2, 3, 1).reshape(N * C, T, CH)
    y, _ = lstm(sd, p + "inter_rnn.rnn.", v, False, None, fast=fast_lstm)
    y = F.linear(y, sd[p + "inter_rnn.proj.weight"], sd[p + "inter_rnn.proj.bias"])
    y = F.layer_norm(y, (CH,), sd[p + "inter_norm.weight"], sd[p + "inter_norm.bias"], 1e-5)
    return x + y.reshape(N, C, T, CH).permute(0, 3, 1, 2)


def _mha_layer(sd: SD, p: str, x: Tensor, nhead: int, causal: bool = False) -> Tensor:
    """MhaSelfAttenLayer.forward (improved=False), lobe/attention.py:187-232, on x [B, L, E]: positional encoding added to
    the attention input only (:209-213, PositionalEncoding :27-34), nn.MultiheadAttention without biases (:49-55), post-norm
    residual blocks.  The layer has a positional encoding iff it holds the ``pos.pe`` buffer."""
    E = x.shape[-1]
    src = x
    if p + "pos.pe" in sd:
        x = x + sd[p + "pos.pe"][: x.size(1), 0].unsqueeze(0)
    mask = None
    if causal:
        L = x.size(1)
        mask = torch.full((L, L), float("-inf")).triu(1)
    x, _ = F.multi_head_attention_forward(
        x.transpose(0, 1), x.transpose(0, 1), x.transpose(0, 1), E, nhead, sd[p + "self_atten.atten.in_proj_weight"], None, None, None,
        False, 0.0, sd[p + "self_atten.atten.out_proj.weight"], None, training=False, need_weights=False, attn_mask=mask)
    x = x.transpose(0, 1)
    x = F.layer_norm(src + x, (E,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    src = x
    x = F.linear(torch.relu(F.linear(x, sd[p + "feedforward.0.weight"], sd[p + "feedforward.0.bias"])),
                 sd[p + "feedforward.3.weight"], sd[p + "feedforward.3.bias"])
    return F.layer_norm(src + x, (E,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)


def _dparn_block2d(sd: SD, p: str, x: Tensor, nhead: int, fast_lstm: bool) -> Tensor:
    """DPARNblock2D.forward, dparn.py:54-108: two transformer encoder layers over the frequency rows of every frame
    (non-causal), Linear -> LayerNorm -> + skip; then the uni-directional inter-chunk LSTM pass of DPRNNblock2D."""
    N, CH, C, T = x.shape
    v = x.permute(0, 3, 2, 1).reshape(N * T, C, CH)
    v = _mha_layer(sd, p + "intra_atten2.", _mha_layer(sd, p + "intra_atten1.", v, nhead), nhead)
    v = F.layer_norm(F.linear(v, sd[p + "intra_fc.weight"], sd[p + "intra_fc.bias"]), (CH,), sd[p + "intra_norm.weight"],
                     sd[p + "intra_norm.bias"], 1e-5)
    x = x + v.reshape(N, T, C, CH).permute(0, 3, 2, 1)
    v = x.permute(0, 2, 3, 1).reshape(N * C, T, CH)
    y, _ = lstm(sd, p + "inter_rnn.rnn.", v, False, None, fast=fast_lstm)
    y = F.linear(y, sd[p + "inter_rnn.proj.weight"], sd[p + "inter_rnn.proj.bias"])
    y = F.layer_norm(y, (CH,), sd[p + "inter_norm.weight"], sd[p + "inter_norm.bias"], 1e-5)
    return x + y.reshape(N, C, T, CH).permute(0, 3, 1, 2)


def dparn(sd: SD, p: str, x: Tensor, a: dict, fast_lstm: bool = True) -> Tensor:
    """DPARN.forward, dparn.py:169-223: the U-Net shell around two DPARNblock2D."""
    h = a["nhead"]
    return _unet_shell(sd, p, x, a, lambda v: _dparn_block2d(sd, p + "dprnn_block2.", _dparn_block2d(sd, p + "dprnn_block1.", v, h, fast_lstm), h, fast_lstm))


def dpcrn(sd: SD, p: str, x: Tensor, a: dict, fast_lstm: bool = True) -> Tensor:
    """DPCRN.forward, dpcrn.py:136-190: the U-Net shell around two DPRNNblock2D."""
    return _unet_shell(sd, p, x, a, lambda v: _dprnn_block2d(sd, p + "dprnn_block2.", _dprnn_block2d(sd, p + "dprnn_block1.", v, fast_lstm), fast_lstm))


def masker_forward(sd: SD, p: str, mcfg: dict, x: Tensor, dvec: Optional[Tensor], fast_lstm: bool = True) -> Tensor:
    if mcfg["type"] == "UnetTcn":
        return unet_tcn(sd, p, x, dvec, mcfg)
    if mcfg["type"] == "DPCRN":
        return dpcrn(sd, p, x, mcfg, fast_lstm)
    if mcfg["type"] == "DPARN":
        return dparn(sd, p, x, mcfg, fast_lstm)
    if mcfg["type"] == "ConvTasNet":
        return conv_tasnet(sd, p, x, dvec, mcfg)
    if mcfg["type"] == "DPRNN":
        return dprnn(sd, p, x, dvec, mcfg, fast_lstm)
    if mcfg["type"] == "SkiM":
        return skim(sd, p, x, dvec, mcfg, fast_lstm)
    raise NotImplementedError(mcfg["type"])


def tse_embedding(sd: SD, cfg: dict, enroll: Tensor) -> Tensor:
    """SoTaskWrapModule.inference_tse_embedding, base_nn.py:724-738 (returns the
    un-squeezed speaker-net output, as the reference does)."""
    if cfg.get("encoder_spk") is not None:
        f = _encode(sd, "encoder_spk.", cfg["encoder_spk"], enroll, cfg["drop_first_bin"])
    else:
        f = _encode(sd, "encoder.", cfg["encoder"], enroll, cfg["drop_first_bin"])
    return speaker_net(sd, "speaker_net.", cfg["speaker_net"], f).unsqueeze(-1)


def inference(
    sd: SD, cfg: dict, noisy: Tensor, enroll: Optional[Tensor] = None, pre_clamp: bool = False, fast_lstm: bool = True
) -> Tensor:
    """SoTaskWrapModule.inference, base_nn.py:690-722.  ``pre_clamp`` returns the
    waveform before _wav_output_constrain (parity is also checked
    there because the clamp can hide errors)."""
    with torch.no_grad():
        feats = _encode(sd, "encoder.", cfg["encoder"], noisy, cfg["drop_first_bin"])
        dvec = None
        if enroll is not None:
            if cfg.get("encoder_spk") is not None:
                dvec = _encode(sd, "encoder_spk.", cfg["encoder_spk"], enroll, cfg["drop_first_bin"])
            else:
                dvec = _encode(sd, "encoder.", cfg["encoder"], enroll, cfg["drop_first_bin"])
            if not cfg["embedding_free_tse"]:
                dvec = speaker_net(sd, "speaker_net.", cfg["speaker_net"], dvec)
        mask = masker_forward(sd, "masker.", cfg["masker"], feats, dvec, fast_lstm)
        mask = get_mask(mask, cfg["mask_constraint"])
        enh = apply_tf_masks(feats, mask, cfg["mask_type"], cfg["f_type"])
        wav = _decode(sd, "encoder.", cfg["encoder"], enh, cfg["drop_first_bin"])
        return wav if pre_clamp else wav_constrain(wav, cfg["output_constraint"])


# --------------------------------------------------------------------------- #
# acceptance metric                     puresound/nnet/loss/sdr.py:263-299
# --------------------------------------------------------------------------- #
def sdr_score(s1: Tensor, s2: Tensor, scaled: bool = True, scale_dependent: bool = False, zero_mean: bool = True,
              sdr_max: Optional[float] = None, eps: float = 1e-8) -> Tensor:
    """SDRLoss.forward without the final sign flip / reduction, loss/sdr.py:139-166 (per item, in dB): [N, L] -> [N, 1]."""
    if zero_mean:
        s1 = s1 - s1.mean(dim=-1, keepdim=True)
        s2 = s2 - s2.mean(dim=-1, keepdim=True)
    l2 = lambda a, b: torch.sum(a * b, -1, keepdim=True)
    s_target = l2(s1, s2) / (l2(s2, s2) + eps) * s2 if scaled else s2
    e_noise = s1 - s2 if scale_dependent else s1 - s_target
    target_norm, noise_norm = l2(s_target, s_target), l2(e_noise, e_noise)
    if sdr_max is not None:
        noise_norm = noise_norm + 10 ** (-sdr_max / 10) * target_norm
    return 10 * torch.log10(target_norm / (noise_norm + eps) + eps)


def si_snr(est: Tensor, ref: Tensor, eps: float = 1e-8) -> Tensor:
    """si_snr(reduction=False): zero-mean, project, 10*log10 power ratio."""
    est = est - est.mean(-1, keepdim=True)
    ref = ref - ref.mean(-1, keepdim=True)
    dot = (est * ref).sum(-1, keepdim=True)
    tgt = dot / ((ref * ref).sum(-1, keepdim=True) + eps) * ref
    err = est - tgt
    return 10 * torch.log10((tgt * tgt).sum(-1) / ((err * err).sum(-1) + eps) + eps)
