"""DPARN on the B200 engine (drop-in for ``puresound.nnet.dparn.DPARNblock2D`` / ``DPARN`` and the attention lobes they use,
reference dparn.py:12-246, lobe/attention.py:8-232; SURVEY.md 8f rank 3, second half - the masker of the ``egs/ns`` recipes
``ns_dparn_v0[_causal]``).  DPCRN with the intra-chunk LSTM replaced by two post-norm transformer encoder layers over the
frequency rows of every frame:

    qkv   = (x + pe) W_in^T              one GEMM; the positional encoding enters as a per-row bias  pe W_in^T
    a     = softmax(q k^T / sqrt(dh)) v  ps_attention (per sequence and head; K, V of a head in shared memory)
    x1    = LayerNorm(x + a W_out^T)     GEMM with the residual epilogue, then the row-norm kernel
    x2    = LayerNorm(x1 + W_2 relu(W_1 x1 + b_1) + b_2)      two GEMMs (ReLU / residual epilogues) + row-norm

then ``Linear -> LayerNorm -> + skip`` (one GEMM epilogue) and the inter-chunk LSTM pass of DPCRN.  The U-Net shell is
``nnet/unet.py``.  Same constructors, sub-module names and ``state_dict`` keys as the reference (``nn.MultiheadAttention`` and
the ``pe`` buffer are kept as parameter holders).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn as nn

from .. import ops
from ._fuse import ParamCache
from .dpcrn import SingleRNN
from .dprnn import dual_path_pass, proj_ln_residual
from .unet import Unet


class PositionalEncoding(nn.Module):
    """Holder of the sin/cos table under the reference's buffer key ``pe`` (lobe/attention.py:8-34)."""

    def __init__(self, d_model: int, dropout: float = 0.1, max_len: int = 5000):
        super().__init__()
        if d_model % 2 != 0:
            raise ValueError(f"Cannot use sin/cos positional encoding with odd dim (got dim={d_model})")
        self.dropout = nn.Dropout(p=dropout)
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, 1, d_model)
        pe[:, 0, 0::2] = torch.sin(position * div_term)
        pe[:, 0, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)


class MHA(nn.Module):
    """Holder with the reference's keys (lobe/attention.py:37-55): ``atten.in_proj_weight``, ``atten.out_proj.weight``."""

    def __init__(self, embed_dim: int, heads: int = 1):
        super().__init__()
        self.atten = nn.MultiheadAttention(embed_dim=embed_dim, num_heads=heads, dropout=0, batch_first=True, bias=False)


class MhaSelfAttenLayer(nn.Module):
    """reference: lobe/attention.py:116-232 (the post-norm encoder layer; ``improved`` = LSTM feed-forward is not used by
    DPARN and raises)."""

    def __init__(self, feats_dim: int, hidden_dim: int, nhead: int, dropout: float = 0.0, improved: bool = False,
                 bidirectional: bool = False, position_encoding: bool = True):
        super().__init__()
        if improved:
            raise NotImplementedError("the improved (LSTM feed-forward) transformer layer is not used by the reference's recipes")
        self.improved, self.bidirectional, self.position_encoding = improved, bidirectional, position_encoding
        self.nhead = nhead
        self.self_atten = MHA(feats_dim, heads=nhead)
        self.self_atten_dropout = nn.Dropout(p=dropout)
        self.norm1 = nn.LayerNorm(feats_dim)
        if position_encoding:
            self.pos = PositionalEncoding(d_model=feats_dim, dropout=dropout)
        self.feedforward = nn.Sequential(nn.Linear(feats_dim, hidden_dim), nn.ReLU(), nn.Dropout(p=dropout),
                                         nn.Linear(hidden_dim, feats_dim), nn.Dropout(p=dropout))
        self.norm2 = nn.LayerNorm(feats_dim)
        self._cache = ParamCache()

    def _pk(self, tag: str, w: torch.Tensor):
        return self._cache.get(tag, [w], lambda: ops.pack_weights(w, w.shape[0], w.shape[1], w.shape[1]))

    def forward_cl(self, x: torch.Tensor, causal: bool = False) -> torch.Tensor:
        """x [B, L, E] (B sequences of L positions) -> [B, L, E]."""
        if self.training and self.self_atten_dropout.p > 0:
            raise NotImplementedError("dropout > 0 in train mode is a training feature (out of scope)")
        B, L, E = x.shape
        att = self.self_atten.atten
        w_in, w_out = att.in_proj_weight, att.out_proj.weight
        # q = k = v = x + pe[position] (lobe/attention.py:209-213; the skip uses x without it): W_in (x + pe) = W_in x + W_in pe,
        # the second term is a [L, 3E] table added through the GEMM's residual input.  Two sequences per GEMM batch item keep
        # the 128-row tiles of the tensor-core kernel full at L = 64.
        G = 2 if (B % 2 == 0 and L <= 64) else 1
        pew = None
        if self.position_encoding:
            def build():
                t, _ = ops.linear(self.pos.pe[:L, 0, :].contiguous().unsqueeze(0), w_in, backend=ops.GEMM_SIMT)
                return t.view(L, 3 * E).repeat(G, 1).contiguous()
            pew = self._cache.get(f"pew{L}x{G}", [w_in, self.pos.pe], build)
        xg = x.reshape(B // G, G * L, E)
        qkv, _ = ops.linear(xg, w_in, w_packed=self._pk("in", w_in), residual=pew, res_strides=(0, 3 * E) if pew is not None else None)
        a = ops.attention(qkv.view(B, L, 3 * E), self.nhead, causal)
        flat = x.reshape(1, B * L, E)
        y, _ = ops.linear(a.view(1, B * L, E), w_out, w_packed=self._pk("out", w_out), residual=flat)
        x1 = ops.rownorm(y, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        f1, f2 = self.feedforward[0], self.feedforward[3]
        h, _ = ops.linear(x1, f1.weight, bias=f1.bias, epi_act=ops.ACT_RELU, w_packed=self._pk("ff1", f1.weight))
        y, _ = ops.linear(h, f2.weight, bias=f2.bias, w_packed=self._pk("ff2", f2.weight), residual=x1)
        return ops.rownorm(y, self.norm2.weight, self.norm2.bias, self.norm2.eps).view(B, L, E)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, causal: bool = False, context_range=None, return_atten_weight: bool = False):
        """x [N, C, T] -> [N, C, T]"""
        if context_range is not None or return_atten_weight:
            raise NotImplementedError("context_range / attention weights are not used on the inference path")
        return ops.transpose(self.forward_cl(ops.transpose(x), causal))


class DPARNblock2D(nn.Module):
    """reference: dparn.py:12-108."""

    def __init__(self, input_size: int, hidden_size: int, nhead: int, dropout: float = 0.0) -> None:
        super().__init__()
        self.intra_atten1 = MhaSelfAttenLayer(input_size, hidden_size, nhead=nhead, dropout=dropout, improved=False,
                                              bidirectional=False, position_encoding=True)
        self.intra_atten2 = MhaSelfAttenLayer(input_size, hidden_size, nhead=nhead, dropout=dropout, improved=False,
                                              bidirectional=False, position_encoding=False)
        self.intra_fc = nn.Linear(input_size, input_size)
        self.intra_norm = nn.LayerNorm(input_size)
        self.inter_rnn = SingleRNN("LSTM", input_size, hidden_size, bidirectional=False, dropout=dropout)
        self.inter_norm = nn.LayerNorm(input_size)
        self._cache = ParamCache()

    def forward_cl(self, x: torch.Tensor) -> torch.Tensor:
        """x [N, T, F, C] -> [N, T, F, C] (both skips on)."""
        N, T, F_, C_ = x.shape
        v = self.intra_atten2.forward_cl(self.intra_atten1.forward_cl(x.view(N * T, F_, C_)))
        fc = self.intra_fc
        P = N * T * F_
        x = proj_ln_residual(self._cache, "fc", v.view(1, P, C_), fc, self.intra_norm, x.reshape(1, P, C_))
        x, _ = dual_path_pass(self._cache, x.view(N, T, F_, C_), self.inter_rnn.rnn, self.inter_rnn.proj, self.inter_norm, "inter", True)
        return x

    @torch.no_grad()
    def forward(self, x: torch.Tensor, intra_skip: bool = True, inter_skip: bool = True) -> torch.Tensor:
        """x [N, ch, C, T] -> [N, ch, C, T]"""
        if not (intra_skip and inter_skip):
            raise NotImplementedError("DPARN always uses both skips")
        return self.forward_cl(x.permute(0, 3, 2, 1).contiguous()).permute(0, 3, 2, 1).contiguous()


class DPARN(Unet):
    """reference: dparn.py:110-246."""

    def __init__(
        self,
        input_type: str = "RI",
        input_dim: int = 512,
        activation_type: str = "PReLU",
        norm_type: str = "bN2d",
        dropout: float = 0.05,
        channels: Tuple = (1, 32, 32, 32, 64, 128),
        transpose_t_size: int = 2,
        transpose_delay: bool = False,
        skip_conv: bool = False,
        kernel_t: Tuple = (2, 2, 2, 2, 2),
        stride_t: Tuple = (1, 1, 1, 1, 1),
        dilation_t: Tuple = (1, 1, 1, 1, 1),
        kernel_f: Tuple = (5, 3, 3, 3, 3),
        stride_f: Tuple = (2, 2, 1, 1, 1),
        dilation_f: Tuple = (1, 1, 1, 1, 1),
        delay: Tuple = (0, 0, 0, 0, 0),
        rnn_hidden: int = 128,
        nhead: int = 1,
        spectral_compress: bool = False,
    ):
        super().__init__(input_type, input_dim, activation_type, norm_type, dropout, channels, transpose_t_size, skip_conv,
                         kernel_t, stride_t, dilation_t, kernel_f, stride_f, dilation_f, delay)
        self.transpose_delay, self.rnn_hidden, self.nhead, self.spectral_compress = transpose_delay, rnn_hidden, nhead, spectral_compress
        if spectral_compress:
            raise NotImplementedError("spectral_compress is not used by the reference's recipes")
        self.dprnn_block1 = DPARNblock2D(input_size=channels[-1], hidden_size=rnn_hidden, nhead=nhead, dropout=dropout)
        self.dprnn_block2 = DPARNblock2D(input_size=channels[-1], hidden_size=rnn_hidden, nhead=nhead, dropout=dropout)

    def _bottleneck(self, x: torch.Tensor, N: int, T: int, dvec) -> torch.Tensor:
        F_, C_ = x.shape[1], x.shape[2]
        y = self.dprnn_block2.forward_cl(self.dprnn_block1.forward_cl(x.view(N, T, F_, C_)))
        return y.reshape(N * T, F_, C_)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [N, C, T] -> [N, C, T]  (reference dparn.py:169-223)."""
        return ops.transpose(self.forward_cl(ops.transpose(x)))

    @property
    def get_args(self) -> Dict:
        a = super().get_args
        a.pop("multi_output")
        a.update({"transpose_delay": self.transpose_delay, "rnn_hidden": self.rnn_hidden, "nhead": self.nhead})
        return a
