"""Skipping-memory LSTM on the B200 engine (drop-in for ``puresound.nnet.skim``: ``MemLSTM``, ``SegLSTM``, ``SkiM``;
reference skim.py:11-469).  First of the "next" rows of the scope table: it is built from the
kernels of the DPRNN path — input projections and ``Linear -> LayerNorm -> + residual`` on ``ps_gemm``, the recurrence on
``ps_lstm`` (tensor-core kernel when hidden_size == 128, exact-fp32 kernel otherwise), FiLM conditioning, segmentation and
the PReLU + 1x1 output conv — so every arithmetic step still runs in this package's CUDA kernels.

Same constructors, sub-module names and construction order as the reference, hence the same ``state_dict`` keys and the
same seeded initialisation.  The segment tensor stays frames-major ``[N*S, K, C]``; the small ``[D, N*S, H]`` memory
tensors are re-laid with torch views (they are K times smaller than the activations).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_PRELU, PRO_AFFINE, Prologue
from ._fuse import ParamCache, prelu_slope
from .dprnn import proj_ln_residual
from .lobe.trivial import FiLM, overlap_geometry


TC_MIN_POSITIONS = 4096  # sequences x steps from which the tensor-core recurrence pays (it loads 256 KB of weights per CTA)


def _lstm_weights(cache: ParamCache, tag: str, rnn: nn.LSTM, use_tc: bool = True):
    """(W_ih stacked over directions, b_ih + b_hh, W_hh^T [D, H, 4H], resident tensor-core image or None, tcgen05 image of
    W_ih or None); the projection rows are permuted to [dir][unit][gate] when the tensor-core recurrence is used."""
    bi = rnn.bidirectional
    H = rnn.hidden_size
    sfx = ["", "_reverse"] if bi else [""]
    srcs = [getattr(rnn, f"{n}_l0{s}") for s in sfx for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]

    def build():
        w_ih = torch.cat([getattr(rnn, f"weight_ih_l0{s}") for s in sfx], 0).contiguous()
        b = torch.cat([getattr(rnn, f"bias_ih_l0{s}") + getattr(rnn, f"bias_hh_l0{s}") for s in sfx], 0).contiguous()
        w_hh_t = torch.stack([getattr(rnn, f"weight_hh_l0{s}").t().contiguous() for s in sfx], 0).contiguous()
        # sizes the tensor-core kernel does not serve (H = 256 of the recipes) always take the CUDA-core kernel's packed image;
        # for the others the tensor-core image only pays from TC_MIN_POSITIONS on
        tc_size = 32 <= H <= 128 and H % 32 == 0
        w_hh_pk = ops.lstm_pack_weights(w_hh_t, H, len(sfx)) if (use_tc or not tc_size) else None
        if w_hh_pk is not None:
            D = len(sfx)
            w_ih = w_ih.view(D, 4, H, -1).permute(0, 2, 1, 3).reshape(D * 4 * H, -1).contiguous()
            b = b.view(D, 4, H).permute(0, 2, 1).reshape(D * 4 * H).contiguous()
        w_ih_pk = ops.pack_weights(w_ih, w_ih.shape[0], w_ih.shape[1], w_ih.shape[1])
        return w_ih, b, w_hh_t, w_hh_pk, w_ih_pk

    return cache.get(tag + ("_tc" if use_tc else "_fp32"), srcs, build)


def _lstm_proj_norm(cache: ParamCache, tag: str, x: torch.Tensor, rnn: nn.LSTM, proj: nn.Linear, norm: nn.LayerNorm,
                    init: Optional[Tuple[torch.Tensor, torch.Tensor]]):
    """x [B, L, C] -> (x + LayerNorm(Linear(LSTM(x, init))), (h_n, c_n) [D, B, H]): one input-projection GEMM, one recurrent
    kernel, one GEMM with the LayerNorm + residual epilogue."""
    B, L, Cn = x.shape
    H, D = rnn.hidden_size, (2 if rnn.bidirectional else 1)
    P = B * L
    w_ih, b, w_hh_t, w_hh_pk, w_ih_pk = _lstm_weights(cache, tag, rnn, use_tc=P >= TC_MIN_POSITIONS)
    gx, _ = ops.linear(x.reshape(1, P, Cn), w_ih, bias=b, w_packed=w_ih_pk)
    h0 = c0 = None
    if init is not None:
        h0, c0 = init[0].contiguous(), init[1].contiguous()
    hseq, state = ops.lstm(gx.view(P, D * 4 * H), w_hh_t, n_seq=B, L=L, H=H, D=D, inner=1, outer_stride=L, inner_stride=0,
                           step_stride=1, h0=h0, c0=c0, want_state=True, w_packed=w_hh_pk, gx_interleaved=w_hh_pk is not None)
    y = proj_ln_residual(cache, tag + "_proj", hseq.view(1, P, D * H), proj, norm, x.reshape(1, P, Cn))
    return y.view(B, L, Cn), state


class MemLSTM(nn.Module):
    """reference: skim.py:11-171 (offline forward; hidden / cell memories of all segments -> next SegLSTM's states)."""

    def __init__(self, hidden_size: int, causal: bool = True, dropout: float = 0.0):
        super().__init__()
        self.hidden_size = hidden_size
        self.causal = causal
        self.input_size = hidden_size if causal else 2 * hidden_size
        self.bi_direct = not causal
        self.h_net = nn.LSTM(self.input_size, self.hidden_size, num_layers=1, bidirectional=self.bi_direct, batch_first=True)
        self.h_dropout = nn.Dropout(p=dropout)
        self.h_proj = nn.Linear(self.hidden_size * (int(self.bi_direct) + 1), self.input_size)
        self.h_norm = nn.LayerNorm(self.input_size)
        self.c_net = nn.LSTM(self.input_size, self.hidden_size, num_layers=1, bidirectional=self.bi_direct, batch_first=True)
        self.c_dropout = nn.Dropout(p=dropout)
        self.c_proj = nn.Linear(self.hidden_size * (int(self.bi_direct) + 1), self.input_size)
        self.c_norm = nn.LayerNorm(self.input_size)
        self._cache = ParamCache()

    @torch.no_grad()
    def forward(self, h: torch.Tensor, c: torch.Tensor, h_states=None, c_states=None, return_all: bool = False, streaming: bool = False):
        """h, c [N, S, D, H] -> ([D, NS, H], [D, NS, H]) (+ the memory LSTMs' own final states with return_all)."""
        if self.training and (self.h_dropout.p > 0 or self.c_dropout.p > 0):
            raise NotImplementedError("dropout > 0 in train mode is a training feature (out of scope)")
        N, S, D, H = h.shape
        outs, finals = [], []
        for tag, v, rnn, proj, norm, st in (("h", h, self.h_net, self.h_proj, self.h_norm, h_states),
                                            ("c", c, self.c_net, self.c_proj, self.c_norm, c_states)):
            v = v.reshape(N, S, D * H).contiguous()
            v, fin = _lstm_proj_norm(self._cache, tag, v, rnn, proj, norm, st)
            v = v.reshape(N * S, D, H).transpose(1, 0).contiguous()  # [D, NS, H]
            if self.causal and not streaming:
                # causal: segment s starts from the memory of segment s-1, the first from zeros (skim.py:105-112)
                z = torch.zeros_like(v)
                z[:, 1:, :] = v[:, :-1, :]
                v = z
            outs.append(v)
            finals.append(fin)
        if return_all:
            return outs[0], outs[1], finals[0], finals[1]
        return outs[0], outs[1]


class SegLSTM(nn.Module):
    """reference: skim.py:174-262."""

    def __init__(self, input_size: int, hidden_size: int, causal: bool = True, dropout: float = 0.0):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.bi_direct = not causal
        self.causal = causal
        self.lstm = nn.LSTM(input_size, hidden_size, num_layers=1, bidirectional=self.bi_direct, batch_first=True)
        self.drop = nn.Dropout(p=dropout)
        self.proj = nn.Linear(hidden_size * (int(self.bi_direct) + 1), input_size)
        self.norm = nn.LayerNorm(input_size)
        self._cache = ParamCache()

    @torch.no_grad()
    def forward(self, x: torch.Tensor, h: Optional[torch.Tensor], c: Optional[torch.Tensor]):
        """x [NS, K, C], (h, c) [D, NS, H] or None -> (x + LN(proj(LSTM(x))), h_n, c_n)."""
        if self.training and self.drop.p > 0:
            raise NotImplementedError("dropout > 0 in train mode is a training feature (out of scope)")
        init = None
        if h is not None or c is not None:
            D = 2 if self.bi_direct else 1
            z = torch.zeros(D, x.shape[0], self.hidden_size, device=x.device)
            init = (z if h is None else h, z if c is None else c)
        y, (hn, cn) = _lstm_proj_norm(self._cache, "seg", x.contiguous(), self.lstm, self.proj, self.norm, init)
        return y, hn, cn


class SkiM(nn.Module):
    """reference: skim.py:265-469."""

    def __init__(
        self,
        input_size: int,
        hidden_size: int,
        output_size: int,
        n_blocks: int = 2,
        seg_size: int = 20,
        seg_overlap: bool = False,
        causal: bool = True,
        embed_dim: int = 0,
        embed_norm: bool = False,
        embed_fusion: Optional[str] = None,
        block_with_embed: Optional[List] = None,
        dropout: float = 0.0,
    ):
        super().__init__()
        self.seg_size = seg_size
        self.seg_overlap = seg_overlap
        self.hidden_size = hidden_size
        self.n_blocks = n_blocks
        self.causal = causal
        self.embed_dim = embed_dim
        self.embed_norm = embed_norm
        self.block_with_embed = block_with_embed
        self.seg_lstm = nn.ModuleList()
        if embed_dim == 0:
            for _ in range(n_blocks):
                self.seg_lstm.append(SegLSTM(input_size, hidden_size, causal=causal, dropout=dropout))
        else:
            self.seg_input_fusion = nn.ModuleList()
            for i in range(n_blocks):
                self.seg_lstm.append(SegLSTM(input_size, hidden_size, causal=causal, dropout=dropout))
                if block_with_embed[i]:
                    if embed_fusion.lower() == "film":
                        self.seg_input_fusion.append(FiLM(input_size, embed_dim, input_norm=True))
                    elif embed_fusion.lower() == "gate":
                        raise NotImplementedError("the Gate fusion (lobe/trivial.py Gate) is not on the path the recipes use")
                    else:
                        raise NameError
                else:
                    self.seg_input_fusion.append(None)
        self.mem_lstm = nn.ModuleList()
        for _ in range(n_blocks - 1):
            self.mem_lstm.append(MemLSTM(hidden_size, causal=causal, dropout=dropout))
        self.output_fc = nn.Sequential(nn.PReLU(), nn.Conv1d(input_size, output_size, 1))
        self._cache = ParamCache()

    def _geometry(self, T: int):
        K = self.seg_size
        if self.seg_overlap:
            _, S = overlap_geometry(T, K)
        else:
            S = (T + (K - T % K)) // K  # always a whole extra segment when T % K == 0 (skim.py:436-440)
        return K, S

    def forward_cl(self, x: torch.Tensor, embed: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [N, T, C] frames-major, embed [N, E] -> [N, T, C_out]."""
        if self.embed_norm and embed is not None:
            embed = ops.l2normalize(embed.contiguous())
        N, T, Cn = x.shape
        K, S = self._geometry(T)
        out = ops.segment(x, K, S, self.seg_overlap).view(N * S, K, Cn)
        h = c = None
        H = self.hidden_size
        for i in range(self.n_blocks):
            if embed is not None and self.block_with_embed[i]:
                out = self.seg_input_fusion[i].forward_cl(out.view(N, S * K, Cn), embed).view(N * S, K, Cn)
            out, h, c = self.seg_lstm[i](out, h, c)
            if i < self.n_blocks - 1:
                h = h.reshape(-1, N, S, H).permute(1, 2, 0, 3)  # [D, NS, H] -> [N, S, D, H]
                c = c.reshape(-1, N, S, H).permute(1, 2, 0, 3)
                h, c = self.mem_lstm[i](h, c)
        merged = ops.merge(out.view(N, S, K, Cn), T, self.seg_overlap)
        fc = self.output_fc[1]
        ones, zeros = self._cache.get("fc_id", [fc.weight], lambda: (torch.ones(Cn, device=x.device), torch.zeros(Cn, device=x.device)))
        fc_pk = self._cache.get("fc_pk", [fc.weight], lambda: ops.pack_weights(fc.weight.view(fc.out_channels, Cn), fc.out_channels, Cn, Cn))
        y, _ = ops.linear(merged, fc.weight.view(fc.out_channels, Cn), w_packed=fc_pk,
                          pro=Prologue(PRO_AFFINE, ACT_PRELU, ones, zeros, 0, None, prelu_slope(self.output_fc[0])), bias=fc.bias)
        return y

    @torch.no_grad()
    def forward(self, x: torch.Tensor, embed: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [N, C, T], embed [N, E] -> [N, C_out, T]"""
        return ops.transpose(self.forward_cl(ops.transpose(x), embed))
