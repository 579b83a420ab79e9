"""Frame-by-frame causal Conv-TasNet on the B200 engine.

The reference ships one streaming wrapper, ``StreamingSkiM``
(puresound/streaming/skim_inference.py:10-252), whose surface is ``init_status()``
and ``step_frame(x, embed) -> [., C, 1]`` with the carried state kept on the
module.  A streaming Conv-TasNet does not exist upstream; this one follows the same
API pattern, and — exactly like the reference tests its streaming model
(test/test_streaming.py:62-116) — its oracle is the *offline* causal forward
``ConvTasNet(causal=True, tcn_norm/dconv_norm in {cLN, bN1d})`` of the same weights.

Differences by design: S concurrent streams advance together (x is ``[S, C, 1]``);
all state lives on the device (one dilation-history ring ``[S, (P-1)*d+1, H]`` per
block plus a step counter), so one hop is a fixed chain of kernels that is captured
once in a CUDA graph and replayed (``StreamingSeparator(use_graph=True)``).

Per block and hop: GEMM(W_in) -> one fused kernel (norm1+PReLU, ring push, causal
dilated depthwise taps from the ring, norm2+PReLU) -> GEMM(W_pw) -> [cLN+PReLU row
kernel | bN1d folded into the next GEMM's prologue] -> GEMM(W_out)+residual.

Speaker conditioning (the only causal Conv-TasNet recipe upstream is a TSE model,
``td_tse_conv_tasnet_v0_causal``, egs/tse/model.py:142-182; embed concat at
conv_tasnet.py:78-83; streaming signature ``step_frame(x, embed)``,
skim_inference.py:177): the embedding of a stream does not change from frame to
frame, so ``cat(x, repeat(e))`` through ``W_in`` is folded ONCE - at
``set_embedding`` / the first ``step_frame`` that sees a new embedding - into a
per-stream bias ``W_in[:, C:] e`` per conditioned block (what the offline path does
per item, nnet/conv_tasnet.py), kept in a static buffer the hop graph reads through
the GEMM's residual input.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib, ops
from ..nnet._fuse import norm_kind, prelu_slope
from ..nnet.conv_tasnet import ConvTasNet
from ..nnet.lobe.encoder import FreeEncDec
from ..ops import ACT_NONE, ACT_PRELU, ACT_RELU, ACT_SIGMOID, PRO_AFFINE, PRO_MASK, Prologue


class StreamingConvTasNet(ConvTasNet):
    """Same constructor as ``ConvTasNet``; requires ``causal=True`` and per-frame norms (cLN or bN1d)."""

    TC_MIN_STREAMS = 64  # concurrent streams from which the per-hop 1x1 convs run on tcgen05

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if not self.causal:
            raise AssertionError("streaming needs causal=True")
        if self.tcn_norm not in ("cLN", "bN1d") or self.dconv_norm not in ("cLN", "bN1d") or self.tcn_norm != self.dconv_norm:
            raise AssertionError("streaming needs per-frame norms: tcn_norm == dconv_norm in {cLN, bN1d} "
                                 "(a global gLN silently breaks causality, SURVEY.md section 7)")
        self._state = None

    # ------------------------------------------------------------------ state
    @torch.no_grad()
    def init_status(self, n_streams: int = 1):
        """(Re)start ``n_streams`` streams: zero history rings and step counter (reference: skim_inference.py:142-167)."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("StreamingConvTasNet runs on a CUDA device only (no CPU fallback)")
        rings, folded = [], []
        for stack in self.tcn_list:
            for blk in stack:
                rl = (blk.kernel - 1) * blk.dilation + 1
                rings.append(torch.zeros(n_streams, rl, blk.hid_channels, device=dev))
                if self.tcn_norm == "bN1d":
                    dsc = blk.dconv[0]
                    norms = (blk.in_conv[1], dsc.depthwise[1], dsc.pointwise[1])
                    if any(n.training for n in norms):
                        raise NotImplementedError("train-mode BatchNorm: call .eval() first")
                    folded.append([ops.bn_fold(n.weight, n.bias, n.running_mean, n.running_var, n.eps) for n in norms])
                else:
                    folded.append(None)
        # per-stream embedding bias of the conditioned blocks (filled by set_embedding), static so a captured hop reads it
        ebias = [torch.zeros(n_streams, blk.hid_channels, device=dev) if blk.emb_dim else None for stack in self.tcn_list for blk in stack]
        self._state = {"S": n_streams, "rings": rings, "folded": folded, "step": torch.zeros(1, dtype=torch.int64, device=dev),
                       "ebias": ebias, "embed_sig": None}
        self.frames_counter = 0

    @classmethod
    def from_offline(cls, masker: ConvTasNet) -> "StreamingConvTasNet":
        """Streaming view of an offline causal ``ConvTasNet`` (e.g. the masker of ``td_tse_conv_tasnet_v0_causal``): same
        blocks, shared parameters and packed-weight caches."""
        if isinstance(masker, cls):
            return masker
        if masker.tcn_layer.lower() != "normal":
            raise NotImplementedError("streaming serves tcn_layer='normal' blocks")
        new = cls(**masker.get_args)
        new.tcn_list = masker.tcn_list
        return new

    @torch.no_grad()
    def set_embedding(self, embed: torch.Tensor):
        """embed [S, E] (or [S, E, 1]): one speaker embedding per stream.  Folds it into the per-stream bias of every
        conditioned block (L2-normalised first when embed_norm, conv_tasnet.py:348-349); a captured hop graph sees the new
        values on its next replay."""
        st = self._state
        if st is None:
            raise RuntimeError("call init_status(n_streams) first")
        if not any(self.tcn_with_embed):
            raise ValueError("this masker has no conditioned block (tcn_with_embed is all zeros)")
        e = embed.reshape(embed.shape[0], -1).to(device=st["step"].device, dtype=torch.float32).contiguous()
        if e.shape != (st["S"], self.embed_dim):
            raise ValueError(f"expected an embedding of shape {(st['S'], self.embed_dim)}, got {tuple(e.shape)}")
        if self.embed_norm:
            e = ops.l2normalize(e)
        j = 0
        for stack in self.tcn_list:
            for blk in stack:
                if blk.emb_dim:
                    Cc, H, E = blk.in_channels, blk.hid_channels, blk.emb_dim
                    w_in = blk.in_conv[0].weight.view(H, Cc + E)
                    ops.gemm(e, w_in[:, Cc:], batch=1, rows=e.shape[0], M=H, K=E, x_batch_stride=0, x_row_stride=E, w_row_stride=Cc + E,
                             out=st["ebias"][j])
                j += 1
        st["embed_sig"] = (embed.data_ptr(), embed._version, tuple(embed.shape))

    # ------------------------------------------------------------------ one hop
    def _block_step(self, blk, j: int, x: torch.Tensor) -> torch.Tensor:
        st = self._state
        S = st["S"]
        Cc, H, E = blk.in_channels, blk.hid_channels, blk.emb_dim
        dsc = blk.dconv[0]
        n1, n2, n3 = blk.in_conv[1], dsc.depthwise[1], dsc.pointwise[1]
        kind = 0 if self.tcn_norm == "cLN" else 1
        xin = x.view(1, S, Cc)
        ebias = st["ebias"][j]  # [S, H] per-stream W_in[:, C:] e of a conditioned block: rides the GEMM's residual input
        # with enough concurrent streams a hop is a [S x 512 x 512] GEMM worth the tensor cores (3xBF16, as offline);
        # few streams keep the exact-fp32 latency tile (the CTA-pair kernel has ~10 us of fixed cost)
        tc = S >= self.TC_MIN_STREAMS
        pk_in = blk._packed("in", blk.in_conv[0].weight, H, Cc, Cc + E) if tc else None
        pk_pw = blk._packed("pw", dsc.pointwise[0].weight, H, H, H) if tc else None
        pk_out = blk._packed("out", blk.out_conv.weight, Cc, H, H) if tc else None
        u1, _ = ops.linear(xin, blk.in_conv[0].weight.view(H, Cc + E), K=Cc, w_row_stride=Cc + E, w_packed=pk_in,
                           residual=None if ebias is None else ebias.view(1, S, H))
        u2 = torch.empty(S, H, device=x.device, dtype=torch.float32)
        d = _lib.StreamDwDesc()
        d.streams, d.C, d.P, d.dilation = S, H, blk.kernel, blk.dilation
        d.u, d.y, d.ring, d.step = u1.data_ptr(), u2.data_ptr(), st["rings"][j].data_ptr(), st["step"].data_ptr()
        dw = dsc.depthwise[0]
        d.w, d.bias = dw.weight.data_ptr(), dw.bias.data_ptr()
        d.norm_kind = kind
        if kind == 0:
            d.eps = n1.eps
            d.n1_a, d.n1_b, d.n2_a, d.n2_b = n1.gamma.data_ptr(), n1.beta.data_ptr(), n2.gamma.data_ptr(), n2.beta.data_ptr()
        else:
            f = st["folded"][j]
            d.eps = 0.0
            d.n1_a, d.n1_b, d.n2_a, d.n2_b = f[0][0].data_ptr(), f[0][1].data_ptr(), f[1][0].data_ptr(), f[1][1].data_ptr()
        d.slope1, d.slope2 = prelu_slope(blk.in_conv[2]).data_ptr(), prelu_slope(dsc.depthwise[2]).data_ptr()
        _lib.check(_lib.load().ps_stream_dwconv_step(C.byref(d), torch.cuda.current_stream().cuda_stream), "ps_stream_dwconv_step")
        ops._launched()
        pw = dsc.pointwise[0]
        u3, _ = ops.linear(u2.view(1, S, H), pw.weight.view(H, H), bias=pw.bias, w_packed=pk_pw)
        slope3 = prelu_slope(dsc.pointwise[2])
        if kind == 0:
            u3n = ops.rownorm(u3, n3.gamma, n3.beta, n3.eps, act=ACT_PRELU, slope=slope3)
            pro = ops.NO_PRO
        else:
            u3n = u3
            f = st["folded"][j]
            pro = Prologue(PRO_AFFINE, ACT_PRELU, f[2][0], f[2][1], 0, None, slope3)
        y, _ = ops.linear(u3n, blk.out_conv.weight.view(Cc, H), pro=pro, bias=blk.out_conv.bias, residual=xin, w_packed=pk_out)
        return y.view(S, Cc)

    def step_frame_cl(self, x: torch.Tensor, embed: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [S, C] (one frame per stream) -> mask logits [S, C]; advances every stream by one frame."""
        if self._state is None:
            raise RuntimeError("call init_status(n_streams) first")
        if x.shape[0] != self._state["S"]:
            raise ValueError(f"expected {self._state['S']} streams, got {x.shape[0]}")
        if any(self.tcn_with_embed):
            if embed is not None:
                if self._state["embed_sig"] != (embed.data_ptr(), embed._version, tuple(embed.shape)):
                    self.set_embedding(embed)  # a new embedding tensor: fold it once (not per frame)
            elif self._state["embed_sig"] is None:
                raise ValueError("this masker is speaker-conditioned: pass embed or call set_embedding(embed) first")
        elif embed is not None:
            raise ValueError("this masker was built without conditioned blocks")
        j = 0
        for stack in self.tcn_list:
            for blk in stack:
                x = self._block_step(blk, j, x)
                j += 1
        _lib.check(_lib.load().ps_stream_advance(self._state["step"].data_ptr(), torch.cuda.current_stream().cuda_stream), "ps_stream_advance")
        ops._launched()
        return x

    @torch.no_grad()
    def step_frame(self, x: torch.Tensor, embed: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [S, C, 1] -> [S, C, 1] (reference pattern: skim_inference.py:176-218)."""
        ops.require_device()
        self.frames_counter += 1
        return self.step_frame_cl(x.reshape(x.shape[0], -1).contiguous(), embed).unsqueeze(-1)


class StreamingSeparator(nn.Module):
    HOP_MMA_MIN_STREAMS = 32  # concurrent streams from which the hop kernel's GEMM phases run on the tensor cores

    """Waveform-in / waveform-out streaming around a task wrapper whose encoder is a ``FreeEncDec`` and whose
    masker is a ``StreamingConvTasNet`` (caller pattern: egs/tse/demo/utils.py:78-118).

    ``step_wave(chunk[S, hop]) -> [S, hop]``: every call consumes one hop of new samples per stream and emits the
    hop of output whose overlap-add is complete.  Unlike the demo's *averaging* overlap-add (utils.py:121-128) the
    emitted samples are the offline decoder's *sum* (ConvTranspose1d), so the concatenated output equals
    ``SoTaskWrapModule.inference`` on the whole signal, delayed by ``win - hop`` samples of priming.
    """

    def __init__(self, model: nn.Module, use_graph: bool = True, use_hop_kernel: bool = True):
        """use_hop_kernel: run a hop as ONE persistent cooperative kernel (ps_stream_hop: every phase of the stack behind grid
        barriers, exact fp32) instead of the chain of ~125 kernels (A/B: use_hop_kernel=False)."""
        super().__init__()
        if not isinstance(model.encoder, FreeEncDec) or not isinstance(model.masker, ConvTasNet):
            raise NotImplementedError("StreamingSeparator needs FreeEncDec + a causal ConvTasNet")
        if model.embedding_free_tse:
            raise NotImplementedError("embedding-free TSE streams through the DPRNN / SkiM wrappers")
        if model.mask_type.lower() != "real" or model.f_type.lower() != "real":
            raise NotImplementedError
        self.model = model
        # an offline causal masker (td_tse_conv_tasnet_v0_causal from recipes.init_model) gets its streaming view here
        self.masker = StreamingConvTasNet.from_offline(model.masker)
        self.use_graph = use_graph
        self.use_hop_kernel = use_hop_kernel and os.environ.get("PS_STREAM_HOP", "1") != "0"
        self._mask_act = {"linear": ACT_NONE, "relu": ACT_RELU, "sigmoid": ACT_SIGMOID}[model.mask_constraint.lower()]
        self._constraint = {"linear": 1, "sigmoid": 2}[model.output_constraint.lower()]
        self._st = None

    @property
    def hop(self) -> int:
        return self.model.encoder.hop_length

    @property
    def win(self) -> int:
        return self.model.encoder.win_length

    @torch.no_grad()
    def init_status(self, n_streams: int = 1, enroll: Optional[torch.Tensor] = None, embed: Optional[torch.Tensor] = None):
        """(Re)start n_streams streams.  For a TSE model pass either the enrollment waveforms `enroll` [S, Le] (the speaker
        net runs once here: SoTaskWrapModule.inference_tse_embedding, base_nn.py:724-738) or ready embeddings `embed` [S, E]."""
        ops.require_device()
        enc = self.model.encoder
        dev = next(self.model.parameters()).device
        if enc.win_length % enc.hop_length != 0:
            raise NotImplementedError("streaming needs win_length to be a multiple of hop_length")
        S, win, hop = n_streams, enc.win_length, enc.hop_length
        self.masker.init_status(S)
        conditioned = any(self.masker.tcn_with_embed)
        if conditioned:
            if embed is None:
                if enroll is None or self.model.speaker_net is None:
                    raise ValueError("a speaker-conditioned model needs init_status(n_streams, enroll=...) or embed=...")
                if enroll.shape[0] != S:
                    raise ValueError(f"expected {S} enrollment utterances, got {enroll.shape[0]}")
                embed = self.model.inference_tse_embedding(enroll.to(dev)).squeeze(-1)
            self.masker.set_embedding(embed)
        elif enroll is not None or embed is not None:
            raise ValueError("this model has no conditioned block")
        self._st = {
            "S": S, "chunks": 0, "prime": win // hop - 1,
            "hist": torch.zeros(S, max(win - hop, 1), device=dev), "frame": torch.zeros(S, win, device=dev),
            "acc": torch.zeros(S, win, device=dev), "chunk": torch.zeros(S, hop, device=dev), "out": torch.zeros(S, hop, device=dev),
            "graph": None,
        }
        # decoder weight transposed once (cache) before any graph capture
        Nf = enc.decoder.in_channels
        self._w_dec_t = enc._cache.get("dec_t", [enc.decoder.weight], lambda: enc.decoder.weight.view(Nf, win).t().contiguous())
        self._hop = self._build_hop_descriptor() if self.use_hop_kernel else None

    def _build_hop_descriptor(self):
        """Descriptor of the persistent hop kernel (ps_stream_hop): per-block pointer table on the device + scratch rows.
        Returns None when the model is outside what that kernel serves (the kernel chain is used then)."""
        st, mst = self._st, self.masker._state
        enc, mk = self.model.encoder, self.masker
        S, win, hop = st["S"], self.win, self.hop
        Cc, H = mk.input_dim, mk.tcn_dim
        blocks = [b for stack in mk.tcn_list for b in stack]
        if Cc % 4 or H % 4 or win % 4 or H > 2048 or enc.encoder.out_channels != Cc:
            return None
        dev = st["out"].device
        kind = 0 if mk.tcn_norm == "cLN" else 1
        arr = (_lib.StreamHopBlock * len(blocks))()
        keep = []
        # From HOP_MMA_MIN_STREAMS concurrent streams on, the GEMM phases run on the tensor cores (mma.sync, 3xBF16 split,
        # ~2^-17 per product like the offline GEMMs); below, exact-fp32 FFMA.  The split of every weight is made once here.
        mma = S >= self.HOP_MMA_MIN_STREAMS and Cc % 16 == 0 and H % 16 == 0 and win % 16 == 0

        def split(w2d: torch.Tensor):
            if not mma:
                return None
            hi = w2d.detach().to(torch.bfloat16)
            lo = (w2d.detach() - hi.float()).to(torch.bfloat16)
            pk = torch.cat([hi, lo], 0).contiguous()
            keep.append(pk)
            return pk.data_ptr()

        for j, blk in enumerate(blocks):
            dsc = blk.dconv[0]
            n1, n2, n3 = blk.in_conv[1], dsc.depthwise[1], dsc.pointwise[1]
            e = arr[j]
            e.w_in, e.w_in_ld = blk.in_conv[0].weight.data_ptr(), blk.in_channels + blk.emb_dim
            e.ebias = ops._p(mst["ebias"][j])
            if kind == 0:
                trip = [(n.gamma, n.beta) for n in (n1, n2, n3)]
            else:
                trip = mst["folded"][j]
            (e.n1_a, e.n1_b), (e.n2_a, e.n2_b), (e.n3_a, e.n3_b) = [(a.data_ptr(), b.data_ptr()) for a, b in trip]
            e.slope1, e.slope2, e.slope3 = (prelu_slope(blk.in_conv[2]).data_ptr(), prelu_slope(dsc.depthwise[2]).data_ptr(),
                                            prelu_slope(dsc.pointwise[2]).data_ptr())
            dw, pw = dsc.depthwise[0], dsc.pointwise[0]
            e.dw_w, e.dw_b = dw.weight.data_ptr(), ops._p(dw.bias)
            e.w_pw, e.b_pw = pw.weight.data_ptr(), ops._p(pw.bias)
            e.w_out, e.b_out = blk.out_conv.weight.data_ptr(), ops._p(blk.out_conv.bias)
            e.ring, e.P, e.dilation = mst["rings"][j].data_ptr(), blk.kernel, blk.dilation
            e.w_in_p = split(blk.in_conv[0].weight.view(H, blk.in_channels + blk.emb_dim)[:, :Cc])
            e.w_pw_p = split(pw.weight.view(H, H))
            e.w_out_p = split(blk.out_conv.weight.view(Cc, H))
        table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        scratch = {k: torch.zeros(S, n, device=dev) for k, n in (("feats", Cc), ("x", Cc), ("u1", H), ("u2", H), ("u3", H), ("frame_out", win))}
        barrier = torch.zeros(2, dtype=torch.int32, device=dev)
        d = _lib.StreamHopDesc()
        d.streams, d.C, d.H, d.win, d.hop, d.n_blocks = S, Cc, H, win, hop, len(blocks)
        d.norm_kind, d.enc_relu, d.mask_act, d.constraint = kind, int(bool(enc.output_active)), self._mask_act, self._constraint
        d.eps = float(blocks[0].in_conv[1].eps) if kind == 0 else 0.0
        d.w_enc, d.w_dec_t, d.blocks = enc.encoder.weight.data_ptr(), self._w_dec_t.data_ptr(), table.data_ptr()
        d.chunk, d.hist, d.frame, d.frame_out = st["chunk"].data_ptr(), st["hist"].data_ptr(), st["frame"].data_ptr(), scratch["frame_out"].data_ptr()
        d.acc, d.out, d.step = st["acc"].data_ptr(), st["out"].data_ptr(), mst["step"].data_ptr()
        d.feats, d.x, d.u1, d.u2, d.u3 = (scratch[k].data_ptr() for k in ("feats", "x", "u1", "u2", "u3"))
        d.barrier = barrier.data_ptr()
        d.w_enc_p, d.w_dec_p = split(enc.encoder.weight.view(Cc, win)), split(self._w_dec_t)
        return {"desc": d, "keep": (table, scratch, barrier, keep)}

    def _hop_kernel(self):
        _lib.check(_lib.load().ps_stream_hop(C.byref(self._hop["desc"]), torch.cuda.current_stream().cuda_stream), "ps_stream_hop")
        ops._launched()

    def _push(self):
        st = self._st
        lib = _lib.load()
        _lib.check(lib.ps_stream_push(st["chunk"].data_ptr(), st["hist"].data_ptr(), st["frame"].data_ptr(), st["S"], self.win, self.hop,
                                      torch.cuda.current_stream().cuda_stream), "ps_stream_push")
        ops._launched()

    def _push_and_compute(self):
        self._push()
        self._compute()

    def _compute(self):
        """encoder GEMM -> masker step -> mask-apply + decoder GEMM -> overlap-add emit, all on static buffers."""
        st = self._st
        enc = self.model.encoder
        S, win, hop = st["S"], self.win, self.hop
        Nf = enc.encoder.out_channels
        feats, _ = ops.linear(st["frame"].view(1, S, win), enc.encoder.weight.view(Nf, win), epi_act=ACT_RELU if enc.output_active else ACT_NONE)
        mask = self.masker.step_frame_cl(feats.view(S, Nf))
        fr, _ = ops.linear(feats, self._w_dec_t, pro=Prologue(PRO_MASK, self._mask_act, x2=mask.view(1, S, Nf)))
        lib = _lib.load()
        _lib.check(lib.ps_stream_ola(fr.data_ptr(), st["acc"].data_ptr(), st["out"].data_ptr(), S, win, hop, self._constraint,
                                     torch.cuda.current_stream().cuda_stream), "ps_stream_ola")
        ops._launched()

    @torch.no_grad()
    def step_wave(self, chunk: torch.Tensor) -> torch.Tensor:
        """chunk [S, hop] (host or device) -> [S, hop] on the same device."""
        st = self._st
        if st is None:
            raise RuntimeError("call init_status(n_streams) first")
        if tuple(chunk.shape) != (st["S"], self.hop):
            raise ValueError(f"expected a chunk of shape {(st['S'], self.hop)}, got {tuple(chunk.shape)}")
        on_host = not chunk.is_cuda
        st["chunk"].copy_(chunk, non_blocking=True)
        st["chunks"] += 1
        hopk = self._hop is not None
        if st["chunks"] <= st["prime"]:
            self._push()
            out = torch.zeros_like(st["out"])  # the first complete window has not arrived yet
            return out.cpu() if on_host else out
        step = self._hop_kernel if hopk else self._push_and_compute  # (the hop kernel assembles the frame itself)
        first = st["chunks"] == st["prime"] + 1
        if not self.use_graph or first:
            step()  # the first hop runs eagerly: it loads the kernels and fills the weight caches
        else:
            if st["graph"] is None:
                # capture the fixed kernel chain of one hop (capture records, it does not execute) ...
                g = torch.cuda.CUDAGraph()
                dev = st["out"].device
                with torch.cuda.device(dev), torch.cuda.graph(g, stream=torch.cuda.Stream(dev)):  # capture stream on OUR device
                    step()
                st["graph"] = g
            st["graph"].replay()  # ... and replay it for this and every later hop
        out = st["out"].clone()
        return out.cpu() if on_host else out
