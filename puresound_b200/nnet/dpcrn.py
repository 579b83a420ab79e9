"""DPCRN on the B200 engine (drop-in for ``puresound.nnet.dpcrn.DPRNNblock2D`` / ``DPCRN``, reference dpcrn.py:11-213;
SURVEY.md 8f rank 3 - the masker of the ``egs/ns`` recipes ``ns_dpcrn_v0[_causal]``).  The U-Net shell is ``nnet/unet.py``
(2-D convs as framed GEMMs); the bottleneck is two dual-path blocks on the ``[N, T, F', C]`` tensor, which is exactly the
``[N, S, K, C]`` layout of the DPRNN kernels with S = frames and K = frequency rows: the intra pass is a bidirectional LSTM
over frequency for every frame, the inter pass a uni-directional LSTM over time for every frequency row (addressed in place),
each followed by ``Linear -> LayerNorm -> + skip`` in one GEMM epilogue (rnn_hidden = 128: tensor-core recurrence).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn

from ._fuse import ParamCache
from .dprnn import dual_path_pass
from .lobe.rnn import SingleRNN  # noqa: F401  (parameter holder of the blocks below; re-exported)
from .unet import Unet


class DPRNNblock2D(nn.Module):
    """reference: dpcrn.py:11-81."""

    def __init__(self, input_size: int, hidden_size: int, dropout: float = 0.0) -> None:
        super().__init__()
        self.intra_rnn = SingleRNN("LSTM", input_size, hidden_size, bidirectional=True, dropout=dropout)
        self.intra_norm = nn.LayerNorm(input_size)
        self.inter_rnn = SingleRNN("LSTM", input_size, hidden_size, bidirectional=False, dropout=dropout)
        self.inter_norm = nn.LayerNorm(input_size)
        self._cache = ParamCache()

    def forward_cl(self, x: torch.Tensor) -> torch.Tensor:
        """x [N, T, F, C] -> [N, T, F, C] (both skips on, as DPCRN calls it)."""
        if self.training and (self.intra_rnn.drop.p > 0 or self.inter_rnn.drop.p > 0):
            raise NotImplementedError("dropout > 0 in train mode is a training feature (out of scope)")
        x, _ = dual_path_pass(self._cache, x, self.intra_rnn.rnn, self.intra_rnn.proj, self.intra_norm, "intra", False)
        x, _ = dual_path_pass(self._cache, x, self.inter_rnn.rnn, self.inter_rnn.proj, self.inter_norm, "inter", True)
        return x

    @torch.no_grad()
    def forward(self, x: torch.Tensor, intra_skip: bool = True, inter_skip: bool = True) -> torch.Tensor:
        """x [N, ch, C, T] -> [N, ch, C, T]"""
        if not (intra_skip and inter_skip):
            raise NotImplementedError("DPCRN always uses both skips (dpcrn.py:166-167)")
        y = self.forward_cl(x.permute(0, 3, 2, 1).contiguous())
        return y.permute(0, 3, 2, 1).contiguous()


class DPCRN(Unet):
    """reference: dpcrn.py:84-213."""

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
        spectral_compress: bool = False,
    ):
        super().__init__(input_type, input_dim, activation_type, norm_type, dropout, channels, transpose_t_size, skip_conv,
                         kernel_t, stride_t, dilation_t, kernel_f, stride_f, dilation_f, delay)
        self.transpose_delay = transpose_delay
        self.rnn_hidden = rnn_hidden
        self.spectral_compress = spectral_compress
        if spectral_compress:
            raise NotImplementedError("spectral_compress is not used by the reference's recipes")
        self.dprnn_block1 = DPRNNblock2D(input_size=channels[-1], hidden_size=rnn_hidden, dropout=dropout)
        self.dprnn_block2 = DPRNNblock2D(input_size=channels[-1], hidden_size=rnn_hidden, dropout=dropout)

    def _bottleneck(self, x: torch.Tensor, N: int, T: int, dvec) -> torch.Tensor:
        F_, C_ = x.shape[1], x.shape[2]
        y = self.dprnn_block2.forward_cl(self.dprnn_block1.forward_cl(x.view(N, T, F_, C_)))
        return y.reshape(N * T, F_, C_)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [N, C, T] -> [N, C, T]  (reference dpcrn.py:136-190)."""
        from .. import ops

        return ops.transpose(self.forward_cl(ops.transpose(x)))

    @property
    def get_args(self) -> Dict:
        a = super().get_args
        a.pop("multi_output")
        a.update({"transpose_delay": self.transpose_delay, "rnn_hidden": self.rnn_hidden})
        return a
