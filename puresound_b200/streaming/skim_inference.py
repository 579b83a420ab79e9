"""Frame-by-frame SkiM on the B200 engine (drop-in for ``puresound.streaming.skim_inference.StreamingSkiM``,
skim_inference.py:10-252): same constructor, ``init_status`` / ``step_frame`` / ``step_chunk`` / ``update_mem_lstm`` /
``reset_seg_lstm_status`` and the same state lists, with one generalisation — ``init_status(n_streams)`` runs that many
independent streams in one call (the reference is single-stream: n_streams = 1 reproduces it).

Every frame is the kernel chain of the offline model at L = 1: FiLM, the SegLSTM input projection, one recurrent step,
``Linear -> LayerNorm -> + residual``, and the PReLU + 1x1 output conv; every ``seg_size`` frames the memory LSTMs advance by
one segment.  The states live on the device between calls.  The reference's own equivalence test (streaming == offline,
test/test_streaming.py:62-116) is the parity statement; ``tests/test_gpu_streaming.py`` repeats it against the oracle.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .. import ops
from ..nnet.skim import SkiM
from ..ops import ACT_PRELU, PRO_AFFINE, Prologue
from ..nnet._fuse import prelu_slope


class StreamingSkiM(SkiM):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.n_streams = 1
        self.frames_counter = 0

    # ------------------------------------------------------------------ state (skim_inference.py:142-174)
    @torch.no_grad()
    def init_status(self, n_streams: int = 1):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("StreamingSkiM runs on a CUDA device only (no CPU fallback)")
        D = int(not self.causal) + 1
        S, H = n_streams, self.hidden_size
        self.n_streams = S
        self.frames_counter = 0
        z = lambda: torch.zeros(D, S, H, device=dev)
        self.seg_lstm_h_states = [z() for _ in range(self.n_blocks)]
        self.seg_lstm_c_states = [z() for _ in range(self.n_blocks)]
        self.mem_lstm_h_hidden = [(z(), z()) for _ in range(self.n_blocks - 1)]
        self.mem_lstm_c_hidden = [(z(), z()) for _ in range(self.n_blocks - 1)]

    @torch.no_grad()
    def reset_seg_lstm_status(self):
        self.seg_lstm_h_states[0] = torch.zeros_like(self.seg_lstm_h_states[0])
        self.seg_lstm_c_states[0] = torch.zeros_like(self.seg_lstm_c_states[0])

    # ------------------------------------------------------------------ one frame
    def _output_fc_cl(self, x: torch.Tensor) -> torch.Tensor:
        """x [S, R, C] -> [S, R, C_out]: PReLU then the 1x1 conv (skim.py:342-344)."""
        Cn = x.shape[-1]
        fc = self.output_fc[1]
        ones, zeros = self._cache.get("fc_id", [fc.weight], lambda: (torch.ones(Cn, device=x.device), torch.zeros(Cn, device=x.device)))
        y, _ = ops.linear(x, fc.weight.view(fc.out_channels, Cn),
                          pro=Prologue(PRO_AFFINE, ACT_PRELU, ones, zeros, 0, None, prelu_slope(self.output_fc[0])), bias=fc.bias)
        return y

    def _frame_cl(self, x: torch.Tensor, embed: Optional[torch.Tensor], h_states: List, c_states: List) -> torch.Tensor:
        """x [S, 1, C]; advances the SegLSTM states in place (lists) and returns the last block's output [S, 1, C]."""
        S, _, Cn = x.shape
        for i in range(self.n_blocks):
            if embed is not None and self.block_with_embed[i]:
                x = self.seg_input_fusion[i].forward_cl(x, embed)
            x, h_states[i], c_states[i] = self.seg_lstm[i](x, h_states[i], c_states[i])
        return x

    @torch.no_grad()
    def step_frame(self, x: torch.Tensor, embed: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [S, 1, C] (the reference: [1, 1, C]), embed [S, E] -> [S, C_out, 1]  (skim_inference.py:176-218)."""
        ops.require_device()
        if not self.causal:
            raise NotImplementedError("frame-by-frame stepping needs causal=True")
        if self.embed_norm and embed is not None:
            embed = ops.l2normalize(embed.contiguous())
        S = self.n_streams
        x = x.reshape(S, 1, -1).contiguous()
        x = self._frame_cl(x, embed, self.seg_lstm_h_states, self.seg_lstm_c_states)
        out = self._output_fc_cl(x)
        self.frames_counter += 1
        if self.frames_counter % self.seg_size == 0:
            self.update_mem_lstm()
            self.reset_seg_lstm_status()
            self.frames_counter = 0
        return out.transpose(1, 2)

    @torch.no_grad()
    def update_mem_lstm(self):
        """End of a segment: every memory LSTM advances one step on the SegLSTM's final states, and its output seeds the
        NEXT block's SegLSTM for the next segment (skim_inference.py:220-252).  Works on the pre-update states."""
        cur_h = [t.clone() for t in self.seg_lstm_h_states]
        cur_c = [t.clone() for t in self.seg_lstm_c_states]
        S, H = self.n_streams, self.hidden_size
        for i in range(self.n_blocks - 1):
            seg_h = cur_h[i].reshape(-1, S, 1, H).permute(1, 2, 0, 3)  # [D, S, H] -> [S, 1, D, H]
            seg_c = cur_c[i].reshape(-1, S, 1, H).permute(1, 2, 0, 3)
            mem_h, mem_c, hid, cell = self.mem_lstm[i](seg_h, seg_c, h_states=self.mem_lstm_h_hidden[i], c_states=self.mem_lstm_c_hidden[i],
                                                       return_all=True, streaming=True)
            self.seg_lstm_h_states[i + 1] = mem_h
            self.seg_lstm_c_states[i + 1] = mem_c
            self.mem_lstm_h_hidden[i] = hid
            self.mem_lstm_c_hidden[i] = cell

    # ------------------------------------------------------------------ one chunk (skim_inference.py:43-139)
    @torch.no_grad()
    def step_chunk(self, x: torch.Tensor, seg_lstm_h_state=None, mem_lstm_h_hidden=None, seg_lstm_c_state=None, mem_lstm_c_hidden=None,
                   embed: Optional[torch.Tensor] = None):
        """x [S, K, C] (one whole segment), explicit state passing as in the reference ->
        (out [S, C_out, K], seg_h [n_blocks-1 x [D,S,H]], mem_h_hidden, seg_c, mem_c_hidden)."""
        ops.require_device()
        if not self.causal:
            raise NotImplementedError("chunk-by-chunk stepping needs causal=True")
        if self.embed_norm and embed is not None:
            embed = ops.l2normalize(embed.contiguous())
        S, K, Cn = x.shape
        nb = self.n_blocks
        if seg_lstm_h_state is not None and seg_lstm_c_state is not None:
            hs = [None] + [seg_lstm_h_state[i] for i in range(nb - 1)]
            cs = [None] + [seg_lstm_c_state[i] for i in range(nb - 1)]
        else:
            hs, cs = [None] * nb, [None] * nb
        if mem_lstm_h_hidden is None and mem_lstm_c_hidden is None:
            mem_lstm_h_hidden, mem_lstm_c_hidden = [None] * (nb - 1), [None] * (nb - 1)
        else:
            mem_lstm_h_hidden, mem_lstm_c_hidden = list(mem_lstm_h_hidden), list(mem_lstm_c_hidden)
        # intra-chunk: the K frames of the segment through every block (a sequence of length K per stream)
        y = x.contiguous()
        for i in range(nb):
            if embed is not None and self.block_with_embed[i]:
                y = self.seg_input_fusion[i].forward_cl(y, embed)
            y, hs[i], cs[i] = self.seg_lstm[i](y, hs[i], cs[i])
        out = self._output_fc_cl(y).transpose(1, 2)
        # inter-segment: one step of every memory LSTM
        H = self.hidden_size
        for i in range(nb - 1):
            seg_h = hs[i].reshape(-1, S, 1, H).permute(1, 2, 0, 3)
            seg_c = cs[i].reshape(-1, S, 1, H).permute(1, 2, 0, 3)
            mem_h, mem_c, hid, cell = self.mem_lstm[i](seg_h, seg_c, h_states=mem_lstm_h_hidden[i], c_states=mem_lstm_c_hidden[i],
                                                       return_all=True, streaming=True)
            hs[i], cs[i] = mem_h, mem_c
            mem_lstm_h_hidden[i], mem_lstm_c_hidden[i] = hid, cell
        return out, hs[:-1], mem_lstm_h_hidden, cs[:-1], mem_lstm_c_hidden
