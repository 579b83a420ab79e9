"""Single-layer recurrent lobe on the B200 engine (drop-in for ``puresound.nnet.lobe.rnn.SingleRNN``, reference
lobe/rnn.py:9-52): ``nn.LSTM`` -> dropout (eval: identity) -> ``nn.Linear`` back to the input width.

It is a parameter holder inside ``DPRNNblock2D`` (DPCRN / DPARN run their LSTMs through ``dual_path_pass``) and a layer of
its own in the speaker net of ``tse_skim_v1_causal`` (egs/tse/model.py:489-499: bidirectional, hidden 192, over the
enrollment's frames).  Engine path: input-projection GEMM -> ``ps_lstm`` (one sequence per item) -> projection GEMM.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from .._fuse import ParamCache


class SingleRNN(nn.Module):
    def __init__(self, rnn_type: str, input_size: int, hidden_size: int, bidirectional: bool = False, dropout: float = 0.0):
        super().__init__()
        rnn_type = rnn_type.upper()
        assert rnn_type in ["RNN", "LSTM", "GRU"], f"Only support 'RNN', 'LSTM' and 'GRU', current type: {rnn_type}"
        if rnn_type != "LSTM":
            raise NotImplementedError("the engine's recurrent kernel is an LSTM (the reference's recipes use LSTM)")
        self.rnn_type, self.input_size, self.hidden_size = rnn_type, input_size, hidden_size
        self.num_direction = int(bidirectional) + 1
        self.rnn = nn.LSTM(input_size, hidden_size, 1, batch_first=True, bidirectional=bidirectional)
        self.drop = nn.Dropout(p=dropout)
        self.proj = nn.Linear(hidden_size * self.num_direction, input_size)
        self._cache = ParamCache()

    def forward_cl(self, x: torch.Tensor) -> torch.Tensor:
        """[N, T, C] frames-major -> [N, T, C]."""
        from ..dprnn import lstm_weights

        N, T, Cn = x.shape
        H, D = self.hidden_size, self.num_direction
        w_ih, b, w_hh_t, w_hh_pk, w_ih_pk = lstm_weights(self._cache, "rnn", self.rnn)
        gx, _ = ops.linear(x.reshape(1, N * T, Cn), w_ih, bias=b, w_packed=w_ih_pk)
        h, _ = ops.lstm(gx.view(N * T, D * 4 * H), w_hh_t, H=H, D=D, n_seq=N, L=T, inner=1, outer_stride=T, inner_stride=0,
                        step_stride=1, w_packed=w_hh_pk, gx_interleaved=w_hh_pk is not None)
        pk = self._cache.get("proj", [self.proj.weight],
                             lambda: ops.pack_weights(self.proj.weight, self.proj.weight.shape[0], self.proj.weight.shape[1], self.proj.weight.shape[1]))
        y, _ = ops.linear(h.view(1, N * T, D * H), self.proj.weight, bias=self.proj.bias, w_packed=pk)
        return y.view(N, T, Cn)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[N, C, T] -> [N, C, T] (reference layout)."""
        return ops.transpose(self.forward_cl(ops.transpose(x)))
