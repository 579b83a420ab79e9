"""Task wrapper on the B200 engine: drop-in for the *inference* surface of
``puresound.nnet.base_nn.SoTaskWrapModule`` (base_nn.py:193-777).

Same constructor signature, attribute names (hence state-dict prefixes
``encoder.``, ``encoder_spk.``, ``masker.``, ``speaker_net.{j}.``) and the same
``inference(noisy, enroll)`` / ``inference_tse_embedding(enroll)`` contracts.  The
whole waveform-in -> waveform-out path stays frames-major on the device: encoder
GEMM -> [speaker net] -> masker -> mask activation+apply (fused into the decoder
GEMM's operand load for real masks) -> synthesis GEMM -> overlap-add with the
output constraint fused.  Training (``forward(**kw)`` -> loss) is out of scope.
"""
from __future__ import annotations

import gc
import os
from collections import OrderedDict
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_NONE, ACT_RELU, ACT_SIGMOID
from .lobe.encoder import ConvEncDec, FbankEnc, FreeEncDec

_MASK_ACT = {"linear": ACT_NONE, "relu": ACT_RELU, "sigmoid": ACT_SIGMOID}
_OUT_CONSTRAINT = {"linear": 1, "sigmoid": 2}


# The mixture's STFT analysis GEMM runs on the tensor cores (3xBF16, ~2^-17 per product) like the masker's GEMMs, whose
# errors reach the iSTFT through the mask anyway: measured at full size (run 81) cfg4 max|dy| 3.0e-5 / pre-clamp 6.6e-5 against
# 2.3e-5 / 5.7e-5 with the exact-fp32 analysis (tolerance 1e-3).  The synthesis GEMM + overlap-add + window-sum-square
# division stay exact fp32.  PS_STFT_EXACT=1 puts the analysis back on the fp32 CUDA-core kernel.
_STFT_EXACT = os.environ.get("PS_STFT_EXACT", "0") == "1"


class SoTaskWrapModule(nn.Module):
    def __init__(
        self,
        encoder: nn.Module,
        masker: nn.Module,
        embedding_free_tse: bool = False,
        encoder_spk: Optional[nn.Module] = None,
        speaker_net: Optional[nn.Module] = None,
        loss_func_wav: Optional[nn.Module] = None,
        loss_func_spk: Optional[nn.Module] = None,
        loss_func_others: Optional[nn.Module] = None,
        f_type: str = "real",
        mask_type: str = "real",
        mask_constraint: str = "linear",
        output_constraint: str = "linear",
        drop_first_bin: bool = False,
        verbose: bool = True,
    ) -> None:
        super().__init__()
        self.f_type, self.mask_type = f_type, mask_type
        self.encoder, self.masker = encoder, masker
        self.embedding_free_tse = embedding_free_tse
        self.encoder_spk, self.speaker_net = encoder_spk, speaker_net
        self.loss_func_wav, self.loss_func_spk, self.loss_func_others = loss_func_wav, loss_func_spk, loss_func_others
        self.mask_constraint, self.output_constraint = mask_constraint, output_constraint
        self.drop_first_bin = drop_first_bin
        # CUDA-graph cache of the waveform -> waveform path, keyed by input shapes (see _run); PS_CUDA_GRAPH=0 disables it
        self.use_cuda_graph = os.environ.get("PS_CUDA_GRAPH", "1") != "0"
        self._graphs = OrderedDict()
        self._host_hooks = None
        self.task = self.check_task()
        if verbose:
            self._verbose()

    # ------------------------------------------------------------------ bookkeeping
    def check_task(self):
        """Task label as the reference derives it (base_nn.py:263-317): 0 SE/BSS, 1 TSE multi-task,
        4 embedding-free TSE, None inference-only TSE."""
        if self.speaker_net is None:
            return 4 if self.embedding_free_tse else 0
        if self.loss_func_wav is None and self.loss_func_spk is None:
            return None
        if self.loss_func_spk is not None and self.loss_func_wav is None:
            return 2
        if self.loss_func_spk is not None and self.loss_func_others is not None:
            return 3
        return 1

    @property
    def overall_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters())

    @property
    def overall_trainable_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def get_state_dict(self):
        return self.state_dict()

    def forward(self, **kwargs):
        raise NotImplementedError("the B200 engine implements inference(); training losses are out of scope (SURVEY.md 8)")

    # ------------------------------------------------------------------ engine path
    def _encode_cl(self, enc: nn.Module, wav: torch.Tensor, exact: bool = True) -> torch.Tensor:
        if isinstance(enc, ConvEncDec):
            return enc.encode_cl(wav, self.drop_first_bin, exact)
        if isinstance(enc, FreeEncDec):
            return enc.encode_cl(wav)
        if isinstance(enc, FbankEnc):  # mel speaker front-end (base_nn.py:373-375: used as it comes)
            return enc.encode_cl(wav, exact)
        raise NotImplementedError(f"encoder {type(enc).__name__} is outside the separator hot path")

    def _speaker_net_cl(self, feats: torch.Tensor) -> torch.Tensor:
        """Apply the speaker net layer by layer (base_nn.py:699-705) on frames-major features -> dvec [N, E]."""
        layers = list(self.speaker_net) if isinstance(self.speaker_net, (nn.ModuleList, nn.Sequential)) else [self.speaker_net]
        x = feats
        for layer in layers:
            if hasattr(layer, "forward_cl"):
                x = layer.forward_cl(x)
            elif isinstance(layer, nn.Conv1d) and layer.kernel_size == (1,) and layer.groups == 1:
                if x.dim() == 2:
                    x = x.unsqueeze(1)
                x, _ = ops.linear(x.contiguous(), layer.weight.view(layer.out_channels, layer.in_channels), bias=layer.bias)
            else:
                raise NotImplementedError(f"speaker-net layer {type(layer).__name__} is outside the separator hot path")
        if x.dim() == 3:
            if x.shape[1] != 1:
                raise ValueError("speaker net must pool over time (output [N, E, 1] in the reference layout)")
            x = x[:, 0, :]
        return x.contiguous()

    def _to_device(self, t: Optional[torch.Tensor]):
        if t is None:
            return None
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("puresound_b200 modules run on a CUDA device only: call model.to('cuda') (no CPU fallback)")
        return t.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()

    def _inference_cl(self, noisy: torch.Tensor, enroll: Optional[torch.Tensor], constrain: bool = True) -> torch.Tensor:
        mask_act = _MASK_ACT.get(self.mask_constraint.lower())
        if mask_act is None:
            raise NotImplementedError
        if constrain:
            constraint = _OUT_CONSTRAINT.get(self.output_constraint.lower())
            if constraint is None:
                raise NameError("Non support type.")
        else:
            constraint = 0
        mt, ft = self.mask_type.lower(), self.f_type.lower()
        if (mt, ft) not in (("real", "real"), ("complex", "complex")):
            if (mt, ft) in (("real", "complex"), ("polar", "polar")):
                raise NotImplementedError("this mask/feature combination is broken upstream (base_nn.py:127, :75)")
            raise NameError

        feats = self._encode_cl(self.encoder, noisy, exact=_STFT_EXACT)
        dvec = None
        if enroll is not None:
            enc = self.encoder if self.encoder_spk is None else self.encoder_spk
            # the enrollment spectrum only feeds the speaker net (never the iSTFT): its analysis GEMM may use tcgen05
            dvec = self._encode_cl(enc, enroll, exact=self.embedding_free_tse)
            if not self.embedding_free_tse:
                dvec = self._speaker_net_cl(dvec)
        mask = self.masker.forward_cl(feats, dvec) if dvec is not None else self.masker.forward_cl(feats)

        if isinstance(self.encoder, ConvEncDec):
            enh = ops.mask_apply(feats, mask, mask_act, mt == "complex")
            return self.encoder.decode_cl(enh, self.drop_first_bin, constraint)
        if mt == "complex":
            enh = ops.mask_apply(feats, mask, mask_act, True)
            return self.encoder.decode_cl(enh, None, ACT_NONE, constraint)
        return self.encoder.decode_cl(feats, mask, mask_act, constraint)

    # ------------------------------------------------------------------ CUDA-graph replay of the whole path
    _GRAPH_SLOTS = 2  # input-shape combinations kept captured (each pins its intermediates in a private pool)

    def _host_prepare(self, device) -> None:
        """Per-call host-side state of modules that have any (SpecAugment draws its random bands from the host RNG): written
        into device buffers BEFORE the forward, so eager runs and graph replays see this call's values."""
        if self._host_hooks is None:
            self._host_hooks = [m for m in self.modules() if m is not self and hasattr(m, "host_prepare")]
        for m in self._host_hooks:
            m.host_prepare(device)

    def _sig_lists(self):
        """Tensors / modules the two signatures walk, collected once: walking the module tree on every call cost 1.9 ms of
        host time per inference() at cfg2's 386 tensors / 417 modules - as long as a whole batch-1 forward, and 8 x that in
        the single-process ShardedSeparator.  (.to() / load_state_dict keep the Parameter objects, so the lists stay valid;
        a model whose sub-modules are replaced after its first inference() must be re-wrapped.)"""
        c = self.__dict__.get("_sig_cache")
        if c is None:
            # SUB-modules only: a list holding `self` would be a reference cycle, the model would then be freed by the cyclic
            # collector at some later allocation - possibly in the middle of ANOTHER model's stream capture, where destroying
            # this one's CUDA graphs / streams invalidates that capture (seen as cudaErrorStreamCaptureInvalidated in the tests)
            mods = [m for m in self.modules() if m is not self]
            c = (list(self.parameters()) + list(self.buffers()), mods, [m for m in mods if isinstance(m, nn.Dropout)])
            self.__dict__["_sig_cache"] = c
        return c

    def _mode_signature(self):
        """train/eval flags of all sub-modules and the dropout probabilities, folded to one hashable value."""
        _, mods, drops = self._sig_lists()
        return (self.training, tuple(m.training for m in mods), tuple(float(m.p) for m in drops))

    def _param_signature(self):
        return tuple([(t.data_ptr(), t._version) for t in self._sig_lists()[0]])

    def _run(self, noisy: torch.Tensor, enroll: Optional[torch.Tensor], constrain: bool = True) -> torch.Tensor:
        """_inference_cl through a captured CUDA graph.  A forward is 170-650 small launches; at batch 1 (cfg1) and for
        the TSE model (cfg4) the launch gaps were longer than the kernels.  The first call with a new combination of input
        shapes runs eagerly (it also builds the packed-weight caches), the second captures, later ones copy the inputs
        into the graph's static buffers and replay.  Any change to a parameter or buffer (load_state_dict, .to(), an
        optimiser step) changes the signature and drops the captured graphs."""
        self._host_prepare(noisy.device)
        if not self.use_cuda_graph:
            return self._inference_cl(noisy, enroll, constrain)
        sig = self._param_signature()
        # everything a replay would otherwise silently ignore is part of the key: shapes, device, the mask / output
        # constraints and the train/eval mode of every sub-module (train-mode BatchNorm / dropout must keep raising)
        key = (tuple(noisy.shape), None if enroll is None else tuple(enroll.shape), constrain, noisy.device.index,
               str(self.mask_constraint).lower(), str(self.output_constraint).lower(), self.f_type, self.mask_type,
               self.drop_first_bin, self._mode_signature())
        ent = self._graphs.get(key)
        if ent is not None and ent["sig"] != sig:
            self._graphs.clear()
            ent = None
        if ent is None:
            self._graphs[key] = {"sig": sig, "graph": None}
            while len(self._graphs) > self._GRAPH_SLOTS:
                self._graphs.popitem(last=False)
            return self._inference_cl(noisy, enroll, constrain)
        self._graphs.move_to_end(key)
        if ent["graph"] is None:
            ent["in"] = noisy.clone()
            ent["enroll"] = None if enroll is None else enroll.clone()
            g = torch.cuda.CUDAGraph()
            # an explicit capture stream ON THE INPUT'S DEVICE: torch.cuda.graph's default capture stream is one class-wide
            # stream created on whichever device captured first, so a replica on another GPU (ShardedSeparator) would
            # capture on a foreign device's stream ("operation not permitted when stream is capturing")
            if not hasattr(self, "_cap_streams"):
                self._cap_streams = {}
            cs = self._cap_streams.get(noisy.device.index)
            if cs is None:
                cs = self._cap_streams[noisy.device.index] = torch.cuda.Stream(noisy.device)
            # no cyclic garbage collection while the stream captures: a dead model of the caller's (reference cycles are
            # enough) that owns CUDA graphs would be finalised wherever the collector happens to run, and destroying a
            # graph inside a capture invalidates it (cudaErrorStreamCaptureInvalidated); torch.cuda.graph no longer
            # collects before capturing
            gc_on = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.device(noisy.device), torch.cuda.graph(g, stream=cs):
                    ent["out"] = self._inference_cl(ent["in"], ent["enroll"], constrain)
            finally:
                if gc_on:
                    gc.enable()
            ent["graph"] = g
            # device tensors the captured kernels read by raw pointer but that live in module-level caches (the iSTFT
            # window-sum-square tables): referenced here so a cache eviction cannot free them under a live graph
            ent["keep"] = [list(m._wsum.values()) for m in self.modules() if hasattr(m, "_wsum")]
        else:
            ent["in"].copy_(noisy, non_blocking=True)
            if enroll is not None:
                ent["enroll"].copy_(enroll, non_blocking=True)
        ent["graph"].replay()
        return ent["out"].clone()

    @staticmethod
    def _to_host(y: torch.Tensor) -> torch.Tensor:
        """Device result -> host tensor through pinned memory (torch's caching host allocator makes the per-call pinned
        buffer cheap): one DMA at PCIe rate instead of the staged pageable copy of ``.cpu()`` (16 MB: 0.7 ms vs 3 ms)."""
        out = torch.empty(y.shape, dtype=y.dtype, pin_memory=True)
        out.copy_(y, non_blocking=True)
        torch.cuda.current_stream(y.device).synchronize()
        return out

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def inference(self, noisy: torch.Tensor, enroll: Optional[torch.Tensor] = None) -> torch.Tensor:
        """noisy [N, L], enroll [N, Le] -> enhanced waveform [N, L'] on the caller's device.

        Host tensors are accepted: they are copied to the model's GPU, and the result is copied back."""
        ops.require_device()
        on_host = not noisy.is_cuda
        y = self._run(self._to_device(noisy), self._to_device(enroll))
        return self._to_host(y) if on_host else y

    @torch.no_grad()
    def inference_stream(self, batches, depth: int = 2, reuse_host_buffers: bool = False):
        """Serving loop over HOST batches: yields the enhanced waveform (pinned host tensor) of every item of `batches` - a
        `noisy` tensor or a `(noisy, enroll)` pair, as `inference` takes them - in order.  Three streams: the host-to-device
        copy of batch i+1 and the device-to-host copy of result i-1 overlap the forward of batch i, so a long run costs
        max(copy, compute) per batch instead of their sum; `depth` results may be in flight before the first is yielded.
        Every batch is computed exactly as `inference` computes it (same kernels, same CUDA-graph replay).
        `reuse_host_buffers=True` takes the results from a ring of depth + 2 pinned buffers instead of a fresh pinned tensor
        per batch (no host allocation in steady state): a yielded tensor is then only valid until depth + 1 further results
        have been yielded - for consumers that use each result before asking for the next ones."""
        from collections import deque

        ops.require_device()
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("puresound_b200 modules run on a CUDA device only: call model.to('cuda') (no CPU fallback)")
        cur = torch.cuda.current_stream(dev)
        h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

        # two persistent device staging slots per input (re-made when a batch changes shape), each guarded by the event of
        # the forward that last read it: no allocation on the copy stream in steady state (blocks freed across streams cannot
        # be reused until their events complete, so per-batch allocations there end in synchronous cudaMalloc calls)
        slots = [{"bufs": [None, None], "free": None} for _ in range(2)]
        state = {"n": 0, "o": 0}
        host_ring = [None] * (depth + 2)

        def host_buffer(shape, dtype):
            if not reuse_host_buffers:
                return torch.empty(shape, dtype=dtype, pin_memory=True)
            i = state["o"] % len(host_ring)
            state["o"] += 1
            if host_ring[i] is None or host_ring[i].shape != shape:
                host_ring[i] = torch.empty(shape, dtype=dtype, pin_memory=True)
            return host_ring[i]

        def upload(item):
            noisy, enroll = item if isinstance(item, (tuple, list)) else (item, None)
            slot = slots[state["n"] % 2]
            state["n"] += 1
            xs = []
            for j, t in enumerate((noisy, enroll)):
                if t is None:
                    xs.append(None)
                    continue
                buf = slot["bufs"][j]
                if buf is None or buf.shape != t.shape:
                    buf = slot["bufs"][j] = torch.empty(t.shape, dtype=torch.float32, device=dev)
                    h2d.wait_stream(cur)  # the new block may still be in use by work queued on the compute stream
                xs.append(buf)
            with torch.cuda.stream(h2d):
                if slot["free"] is not None:
                    h2d.wait_event(slot["free"])
                for buf, t in zip(xs, (noisy, enroll)):
                    if buf is not None:
                        buf.copy_(t, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(h2d)
            return xs[0], xs[1], ev, slot

        it = iter(batches)
        pending = deque()
        nxt = next(it, None)
        up = upload(nxt) if nxt is not None else None
        while up is not None:
            xn, xe, ev, slot = up
            nxt = next(it, None)
            up = upload(nxt) if nxt is not None else None  # the next batch's copy runs under this batch's forward
            cur.wait_event(ev)
            y = self._run(xn, xe)
            done = torch.cuda.Event()
            done.record(cur)
            slot["free"] = done
            with torch.cuda.stream(d2h):
                d2h.wait_event(done)
                out = host_buffer(y.shape, y.dtype)
                out.copy_(y, non_blocking=True)
                y.record_stream(d2h)
                fin = torch.cuda.Event()
                fin.record(d2h)
            pending.append((out, fin))
            if len(pending) > depth:
                o, f = pending.popleft()
                f.synchronize()
                yield o
        while pending:
            o, f = pending.popleft()
            f.synchronize()
            yield o

    @torch.no_grad()
    def inference_pre_constraint(self, noisy: torch.Tensor, enroll: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The waveform before ``_wav_output_constrain`` (parity is also checked here: the clamp hides errors)."""
        ops.require_device()
        on_host = not noisy.is_cuda
        y = self._run(self._to_device(noisy), self._to_device(enroll), constrain=False)
        return self._to_host(y) if on_host else y

    @torch.no_grad()
    def inference_tse_embedding(self, enroll: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Speaker embedding of an enrollment utterance, [N, E, 1] (base_nn.py:724-738)."""
        ops.require_device()
        on_host = not enroll.is_cuda
        enc = self.encoder if self.encoder_spk is None else self.encoder_spk
        enroll = self._to_device(enroll)
        self._host_prepare(enroll.device)
        dvec = self._speaker_net_cl(self._encode_cl(enc, enroll, exact=False)).unsqueeze(-1)
        return dvec.cpu() if on_host else dvec

    def _verbose(self):
        """Same probe as the reference (base_nn.py:740-777): half the input is +inf and the first/last NaN in
        the output gives look-ahead / receptive field, so kernels must propagate NaN/Inf.  Leaves the model in
        train mode like the reference does."""
        import numpy as np

        print("---------------Verbose logging---------------")
        self.eval()
        print(f"Current training mode is: {self.training}")
        print(f"Total params: {self.overall_parameters}")
        if next(self.parameters()).is_cuda:
            x = torch.rand(1, 10 * 16000)
            x[..., 5 * 16000:] = np.inf
            x_spk = torch.rand(1, 10 * 16000)
            needs_enroll = self.speaker_net is not None or self.embedding_free_tse
            y = self.inference(x, x_spk if needs_enroll else None).detach()
            nan_idx = np.where(np.isnan(y.numpy()))[-1]
            lookahead = nan_idx[0]
            print("Lookahead(samples): infinite" if lookahead == 0 else f"Lookahead(samples): {80000 - lookahead}")
            x = torch.rand(1, 10 * 16000)
            x[..., : -5 * 16000] = np.inf
            y = self.inference(x, x_spk if needs_enroll else None).detach()
            receptive = np.where(np.isnan(y.numpy()))[-1][-1]
            if receptive - (80000 - 1) == 80000:
                print("Receptive Fields(samples): infinite")
            else:
                print(f"Receptive Fields(samples): {receptive - (80000 - 1)}")
        else:
            print("(look-ahead / receptive-field probe needs the model on a CUDA device; skipped)")
        self.train()
        print(f"Current training mode is: {self.training}")
        print("---------------Verbose logging---------------")
