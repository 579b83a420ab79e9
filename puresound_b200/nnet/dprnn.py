"""Dual-path RNN on the B200 engine (drop-in for ``puresound.nnet.dprnn.DPRNN``).

Same constructor, sub-module tree (``nn.LSTM`` / ``nn.Linear`` / ``nn.LayerNorm``
holders, so state-dict keys are identical) and ``forward(x[N,C,T], embed)``.

The whole stack works on one frames-major ``[N, S, K, C]`` tensor that never
moves: the intra-chunk pass treats it as ``N*S`` sequences over ``K`` and the
inter-chunk pass as ``N*K`` sequences over ``S`` purely through the LSTM kernel's
strided position function, so the reference's four permute+contiguous copies per
block (dprnn.py:155,165-169,176-178) disappear.  Per pass: one GEMM for the input
projections of both directions, one recurrent kernel, one projection GEMM, one
LayerNorm+residual kernel.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_PRELU, PRO_AFFINE, Prologue
from ._fuse import ParamCache, prelu_slope
from .lobe.trivial import FiLM, overlap_geometry


def lstm_weights(cache: ParamCache, tag: str, rnn: nn.LSTM):
    """(W_ih stacked over directions [D*4H, C], b_ih + b_hh [D*4H], W_hh transposed [D, H, 4H], resident tensor-core image of
    W_hh or None, tcgen05 image of W_ih or None); with the tensor-core recurrence the projection rows are [dir][unit][gate]."""
    sfx = ["", "_reverse"] if rnn.bidirectional else [""]
    H = rnn.hidden_size
    srcs = [getattr(rnn, f"{n}_l0{s}") for s in sfx for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]

    def build():
        w_ih = torch.cat([getattr(rnn, f"weight_ih_l0{s}") for s in sfx], 0).contiguous()
        b = torch.cat([getattr(rnn, f"bias_ih_l0{s}") + getattr(rnn, f"bias_hh_l0{s}") for s in sfx], 0).contiguous()
        w_hh_t = torch.stack([getattr(rnn, f"weight_hh_l0{s}").t().contiguous() for s in sfx], 0).contiguous()
        # resident image for the tensor-core recurrence (None unless H <= 128 and H % 32 == 0), tcgen05 image of W_ih for the projections
        w_hh_pk = ops.lstm_pack_weights(w_hh_t, H, len(sfx))
        if w_hh_pk is not None:
            # tensor-core recurrence: order the projection rows [dir][unit][gate] so the four gates of a unit are one
            # 16-byte load of gx (ps_lstm_t.gx_interleaved)
            D = len(sfx)
            w_ih = w_ih.view(D, 4, H, -1).permute(0, 2, 1, 3).reshape(D * 4 * H, -1).contiguous()
            b = b.view(D, 4, H).permute(0, 2, 1).reshape(D * 4 * H).contiguous()
        w_ih_pk = ops.pack_weights(w_ih, w_ih.shape[0], w_ih.shape[1], w_ih.shape[1])
        return w_ih, b, w_hh_t, w_hh_pk, w_ih_pk

    return cache.get(tag, srcs, build)


def dual_path_pass(cache: ParamCache, out: torch.Tensor, rnn: nn.LSTM, proj: nn.Linear, norm: nn.LayerNorm, tag: str, inter: bool,
                   init=None, want_state: bool = False):
    """One intra- or inter-chunk pass on out [N, S, K, C]: out + LN(Linear(LSTM(out))) (dprnn.py:157-178; the same pattern
    is DPRNNblock2D of DPCRN, dpcrn.py:48-80, with S = frames and K = frequency rows).  Intra: N*S sequences over K; inter: N*K
    sequences over S, addressed in place (no permute)."""
    N, S, K, Cn = out.shape
    H, D = rnn.hidden_size, (2 if rnn.bidirectional else 1)
    w_ih, b, w_hh_t, w_hh_pk, w_ih_pk = lstm_weights(cache, tag, rnn)
    P = N * S * K
    flat = out.view(1, P, Cn)
    gx, _ = ops.linear(flat, w_ih, bias=b, w_packed=w_ih_pk)  # [1, P, D*4H]
    if inter:
        geo = dict(n_seq=N * K, L=S, inner=K, outer_stride=S * K, inner_stride=1, step_stride=K)
    else:
        geo = dict(n_seq=N * S, L=K, inner=1, outer_stride=K, inner_stride=0, step_stride=1)
    h0 = c0 = None
    if init is not None:
        h0, c0 = init[0].contiguous(), init[1].contiguous()
    h, state = ops.lstm(gx.view(P, D * 4 * H), w_hh_t, H=H, D=D, h0=h0, c0=c0, want_state=want_state, w_packed=w_hh_pk,
                        gx_interleaved=w_hh_pk is not None, **geo)
    # Linear -> LayerNorm -> + residual in one kernel (LayerNorm in the GEMM epilogue when Cn == 128; otherwise the
    # library runs the row-norm kernel after the GEMM)
    new = proj_ln_residual(cache, tag + "_proj", h.view(1, P, D * H), proj, norm, out.view(1, P, Cn))
    return new.view(N, S, K, Cn), state


def proj_ln_residual(cache: ParamCache, tag: str, h: torch.Tensor, proj: nn.Linear, norm: nn.LayerNorm, residual: torch.Tensor) -> torch.Tensor:
    """residual + LayerNorm(Linear(h)) on [1, P, *] tensors with the packed weight cached per module."""
    pk = cache.get(tag, [proj.weight], lambda: ops.pack_weights(proj.weight, proj.weight.shape[0], proj.weight.shape[1], proj.weight.shape[1]))
    return ops.linear_ln_residual(h, proj.weight, proj.bias, norm.weight, norm.bias, norm.eps, residual, w_packed=pk)


class DPRNN(nn.Module):
    """reference: dprnn.py:10-244."""

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
        block_with_embed: Optional[List] = None,
        embedding_free_tse: bool = False,
    ):
        super().__init__()
        self.seg_size, self.seg_overlap = seg_size, seg_overlap
        self.input_size, self.hidden_size = input_size, hidden_size
        self.bi_direct = not causal
        self.n_blocks = n_blocks
        self.embed_dim, self.embed_norm = embed_dim, embed_norm
        self.block_with_embed = block_with_embed
        self.embedding_free_tse = embedding_free_tse

        self.input_film = nn.ModuleList()
        self.intra_rnn, self.intra_proj, self.intra_norm = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.inter_rnn, self.inter_norm, self.inter_proj = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        D = 2 if self.bi_direct else 1
        for i in range(n_blocks):
            self.intra_rnn.append(nn.LSTM(input_size, hidden_size, num_layers=1, bidirectional=self.bi_direct, batch_first=True))
            conditioned = embed_dim != 0 and block_with_embed[i]
            self.input_film.append(FiLM(input_size, embed_dim, input_norm=True) if conditioned else None)
            self.intra_proj.append(nn.Linear(hidden_size * D, input_size))
            self.intra_norm.append(nn.LayerNorm(input_size))
            self.inter_rnn.append(nn.LSTM(input_size, hidden_size, num_layers=1, bidirectional=self.bi_direct, batch_first=True))
            self.inter_proj.append(nn.Linear(hidden_size * D, input_size))
            self.inter_norm.append(nn.LayerNorm(input_size))
        self.output_fc = nn.Sequential(nn.PReLU(), nn.Conv1d(input_size, output_size, 1))
        self._cache = ParamCache()

    # ------------------------------------------------------------------ helpers
    def _lstm_weights(self, tag: str, rnn: nn.LSTM):
        return lstm_weights(self._cache, tag, rnn)

    def _geometry(self, T: int):
        K = self.seg_size
        if self.seg_overlap:
            rest, S = overlap_geometry(T, K)
        else:
            rest = K - T % K  # always >= 1: a whole extra segment when T % K == 0 (dprnn.py:141-143)
            S = (T + rest) // K
        return K, S

    def _pass(self, out: torch.Tensor, rnn: nn.LSTM, proj: nn.Linear, norm: nn.LayerNorm, tag: str, inter: bool,
              init=None, want_state: bool = False):
        return dual_path_pass(self._cache, out, rnn, proj, norm, tag, inter, init, want_state)

    def _blocks(self, seg: torch.Tensor, film_embed, inits, collect_hidden: bool):
        out = seg
        hidden = []
        for i in range(self.n_blocks):
            if film_embed is not None and self.block_with_embed is not None and self.block_with_embed[i]:
                N, S, K, Cn = out.shape
                out = self.input_film[i].forward_cl(out.view(N, S * K, Cn), film_embed).view(N, S, K, Cn)
            out, _ = self._pass(out, self.intra_rnn[i], self.intra_proj[i], self.intra_norm[i], f"intra{i}", False)
            out, st = self._pass(out, self.inter_rnn[i], self.inter_proj[i], self.inter_norm[i], f"inter{i}", True,
                                 init=inits[i], want_state=collect_hidden)
            hidden.append(st)
        return hidden if collect_hidden else out

    # ------------------------------------------------------------------ engine path
    def hidden_states_cl(self, x: torch.Tensor):
        """Embedding-free TSE: final (h_n, c_n) of every inter-chunk LSTM on the enrollment features
        (reference dprnn.py:193-244).  x [N, T, C]."""
        K, S = self._geometry(x.shape[1])
        seg = ops.segment(x, K, S, self.seg_overlap)
        return self._blocks(seg, None, [None] * self.n_blocks, True)

    def forward_cl(self, x: torch.Tensor, embed: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [N, T, C]; embed [N, E] (FiLM) or enrollment features [N, Te, C] (embedding-free TSE)."""
        if self.embedding_free_tse:
            assert embed is not None and embed.dim() == 3, "embedding free tse need enrollment waveform as input."
            inits = self.hidden_states_cl(embed)
            film_embed = None
        else:
            inits = [None] * self.n_blocks
            film_embed = embed
            if film_embed is not None and self.embed_norm:
                film_embed = ops.l2normalize(film_embed.contiguous())
        N, T, Cn = x.shape
        K, S = self._geometry(T)
        seg = ops.segment(x, K, S, self.seg_overlap)
        out = self._blocks(seg, film_embed, inits, False)
        merged = ops.merge(out, T, self.seg_overlap)
        ones, zeros = self._cache.get("fc_id", [self.output_fc[1].weight],
                                      lambda: (torch.ones(Cn, device=x.device), torch.zeros(Cn, device=x.device)))
        fc = self.output_fc[1]
        fc_pk = self._cache.get("fc_pk", [fc.weight], lambda: ops.pack_weights(fc.weight.view(fc.out_channels, Cn), fc.out_channels, Cn, Cn))
        y, _ = ops.linear(merged, fc.weight.view(fc.out_channels, Cn), w_packed=fc_pk,
                          pro=Prologue(PRO_AFFINE, ACT_PRELU, ones, zeros, 0, None, prelu_slope(self.output_fc[0])), bias=fc.bias)
        return y

    # ------------------------------------------------------------------ reference-layout API
    @torch.no_grad()
    def forward(self, x: torch.Tensor, embed: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [N, C, T]; embed [N, E] or enrollment features [N, C, Te] -> [N, C_out, T]"""
        if embed is not None and embed.dim() == 3:
            embed = ops.transpose(embed)
        return ops.transpose(self.forward_cl(ops.transpose(x), embed))
