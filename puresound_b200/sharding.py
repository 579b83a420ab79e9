"""Multi-GPU plumbing for a path that needs no data-path collective.

Utterances (and streaming sessions) are independent in eval mode, so the batch is cut into contiguous per-rank
slices.  ``torch.distributed`` is used only to agree on timing: a barrier around the timed region and
a MAX reduction of the per-rank device time.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def shard_slice(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of ``total`` utterances owned by ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """MAX of a per-rank scalar over the default process group (identity when not initialised)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(device: Optional[torch.device] = None) -> None:
    import torch.distributed as dist

    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)


class ShardedSeparator:
    """One process, several GPUs of one box: a host batch is cut into contiguous utterance slices (``shard_slice``), slice g
    runs on replica g of the model on device g, and the enhanced waveforms are gathered, in order, into one pinned host
    tensor.  No collective and no peer traffic: items never mix in eval mode (the reference's analogue is the
    scatter/gather of ``nn.DataParallel``, puresound/task/base.py:226-229, which re-broadcasts the weights every call - here
    every device keeps its replica, its packed-weight caches and its captured CUDA graphs).

    Per device: a persistent pinned staging buffer for its input slice, one side stream, the replica's own graph cache.
    One host thread issues the work of all devices back to back (after the first two calls per shape a forward is one
    graph replay, ~0.1 ms of host time), then waits for every device: wall time = the slowest slice, not the sum.

    Results equal the single-device forward of the whole batch bit for bit as long as a slice and the whole batch are
    served by the same kernel variants (an item's tiles, statistics slots and merge order do not depend on its
    neighbours); the 1x1-conv GEMM picks its tile shape by the number of tiles of a launch (ps_gemm_pair.cu:
    gemm_pair_few_tiles), so a slice of 8 utterances and a batch of 64 can differ in the last bits of the 3xBF16 sums -
    measured (run 17) at 4 and 8 devices, inside the parity tolerances of tests/test_gpu_full.py.

    ``ShardedSeparator(model, devices)`` replicates ``model`` (an engine ``SoTaskWrapModule``); ``runners=`` swaps the
    per-device callables for tests of the split / gather logic on a box without GPUs."""

    def __init__(self, model=None, devices=None, runners=None):
        self._pin = {}
        if runners is not None:
            self.runners, self.devices, self.replicas = list(runners), [None] * len(runners), []
            return
        import copy
        from collections import OrderedDict

        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        if not devices:
            raise RuntimeError("ShardedSeparator needs at least one CUDA device (no CPU fallback)")
        self.devices = [torch.device("cuda", d) if isinstance(d, int) else torch.device(d) for d in devices]
        self.replicas = []
        saved, model._graphs = model._graphs, OrderedDict()  # captured graphs / capture streams belong to the source's device
        saved_cs = model.__dict__.pop("_cap_streams", None)
        model.__dict__.pop("_sig_cache", None)  # (lists of the source's tensors / modules: rebuilt per replica)
        try:
            for dev in self.devices:
                self.replicas.append(copy.deepcopy(model).to(dev).eval())
        finally:
            model._graphs = saved
            if saved_cs is not None:
                model._cap_streams = saved_cs
        self.streams = [torch.cuda.Stream(dev) for dev in self.devices]
        self.runners = [self._device_runner(g) for g in range(len(self.devices))]

    # ------------------------------------------------------------------ per-device work (asynchronous)
    def _staging(self, g: int, tag: str, shape) -> torch.Tensor:
        key = (g, tag)
        buf = self._pin.get(key)
        if buf is None or buf.shape != tuple(shape):
            buf = self._pin[key] = torch.empty(tuple(shape), dtype=torch.float32, pin_memory=True)
        return buf

    def _device_runner(self, g: int):
        dev, rep, stream = self.devices[g], self.replicas[g], self.streams[g]

        def run(noisy: torch.Tensor, enroll: Optional[torch.Tensor], out: torch.Tensor):
            """Issue H2D -> forward -> D2H for this device's slice; returns the event to wait on."""
            with torch.cuda.device(dev), torch.cuda.stream(stream):
                xs = []
                for tag, t in (("noisy", noisy), ("enroll", enroll)):
                    if t is None:
                        xs.append(None)
                        continue
                    st = t if t.is_pinned() else self._staging(g, tag, t.shape).copy_(t)
                    xs.append(st.to(dev, non_blocking=True))
                y = rep._run(xs[0], xs[1])
                out.copy_(y, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
            return ev

        return run

    # ------------------------------------------------------------------ public API (the wrapper's inference contract)
    @torch.no_grad()
    def inference(self, noisy: torch.Tensor, enroll: Optional[torch.Tensor] = None, reuse_output: bool = False) -> torch.Tensor:
        """noisy [N, L] (+ enroll [N, Le]) HOST tensors -> enhanced waveforms [N, L'] (pinned host tensor), item order kept.
        reuse_output=True returns the same pinned buffer for every call with these shapes (serving loops that consume a
        result before asking for the next; a fresh pinned allocation per call costs a device-wide stall now and then)."""
        if noisy.dim() != 2 or (enroll is not None and enroll.shape[0] != noisy.shape[0]):
            raise ValueError("expected noisy [N, L] and enroll [N, Le] with the same N")
        n, world = noisy.shape[0], len(self.runners)
        spans = [shard_slice(n, g, world) for g in range(world)]
        out = None
        pending = []
        for g, (a, b) in enumerate(spans):
            if a == b:
                continue  # fewer utterances than devices
            if out is None:
                # the output length is only known from the model: run the first slice's shape probe lazily
                out = self._out_buffer(n, noisy.shape[1], None if enroll is None else enroll.shape[1], reuse_output)
            pending.append(self.runners[g](noisy[a:b], None if enroll is None else enroll[a:b], out[a:b]))
        for ev in pending:
            if ev is not None:
                ev.synchronize()
        return out

    def _out_buffer(self, n: int, L: int, Le, reuse: bool) -> torch.Tensor:
        key = ("out", n, L, Le)
        buf = self._pin.get(key) if reuse else None
        if buf is None:
            buf = torch.empty(n, self.output_length(L), dtype=torch.float32, pin_memory=torch.cuda.is_available())
            if reuse:
                self._pin[key] = buf
        return buf

    def output_length(self, L: int) -> int:
        """Samples the decoder returns for an L-sample input (FreeEncDec / ConvEncDec: (T-1)*hop + win)."""
        if self.replicas:
            enc = self.replicas[0].encoder
            win = getattr(enc, "win_length", None) or getattr(enc, "n_fft")
            hop = enc.hop_length
            return ((L - win) // hop) * hop + win
        return L
