"""N>1 plumbing on CPU (gloo, world size 2): contiguous utterance shards, no data-path collective, max-over-ranks timing.
The per-rank 'forward' here is the CPU oracle — the point is the host logic, not the kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from puresound_b200.sharding import max_over_ranks, shard_slice


def test_shard_slices_partition_the_batch():
    for total in (1, 2, 7, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [shard_slice(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_slice(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import describe as D
    from oracle import separator_ref as R
    from puresound_b200 import testing
    from puresound_b200.nnet.base_nn import SoTaskWrapModule
    from puresound_b200.nnet.conv_tasnet import ConvTasNet
    from puresound_b200.nnet.lobe.encoder import FreeEncDec

    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)  # weights replicated: every rank builds the same model
    m = SoTaskWrapModule(FreeEncDec(32, 16, 16), ConvTasNet(16, 0, False, tcn_dim=24, per_tcn_stack=2, repeat_tcn=1, tcn_with_embed=[0, 0]),
                         mask_constraint="ReLU", verbose=False).eval()
    wav = testing.white(5, 800, seed=3)  # the global batch (same on every rank), 5 utterances -> shards of 3 and 2
    a, b = shard_slice(wav.shape[0], rank, world)
    y = R.inference(m.state_dict(), D.describe(m), wav[a:b])
    torch.save({"span": (a, b), "y": y}, os.path.join(out_dir, f"rank{rank}.pt"))
    t = max_over_ranks(10.0 + rank)  # pretend per-rank device time
    assert t == 10.0 + world - 1
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_without_collective(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    assert [p["span"] for p in parts] == [(0, 3), (3, 5)]
    # the concatenation of the shards equals the un-sharded forward: items never mix, so no collective is needed
    from oracle import describe as D
    from oracle import separator_ref as R
    from puresound_b200 import testing
    from puresound_b200.nnet.base_nn import SoTaskWrapModule
    from puresound_b200.nnet.conv_tasnet import ConvTasNet
    from puresound_b200.nnet.lobe.encoder import FreeEncDec

    torch.manual_seed(0)
    m = SoTaskWrapModule(FreeEncDec(32, 16, 16), ConvTasNet(16, 0, False, tcn_dim=24, per_tcn_stack=2, repeat_tcn=1, tcn_with_embed=[0, 0]),
                         mask_constraint="ReLU", verbose=False).eval()
    full = R.inference(m.state_dict(), D.describe(m), testing.white(5, 800, seed=3))
    assert torch.allclose(torch.cat([p["y"] for p in parts]), full, atol=1e-6)


def test_sharded_separator_split_and_gather_order():
    """ShardedSeparator's host logic with stub per-device runners (no GPU): contiguous slices, uneven batches, fewer items
    than devices, outputs gathered in item order, enroll sliced with its mixture."""
    from puresound_b200.sharding import ShardedSeparator

    seen = []

    def make(g):
        def run(noisy, enroll, out):
            seen.append((g, noisy.shape[0]))
            out.copy_(noisy * (1 if enroll is None else 2) + (0 if enroll is None else enroll[:, :1]))
            return None
        return run

    sep = ShardedSeparator(runners=[make(g) for g in range(4)])
    x = torch.arange(7 * 5, dtype=torch.float32).view(7, 5)
    assert torch.equal(sep.inference(x), x)
    assert seen == [(0, 2), (1, 2), (2, 2), (3, 1)]
    e = torch.arange(7, dtype=torch.float32).view(7, 1).repeat(1, 3)
    assert torch.equal(sep.inference(x, e), 2 * x + e[:, :1])
    seen.clear()
    assert torch.equal(sep.inference(x[:2]), x[:2])
    assert seen == [(0, 1), (1, 1)]
    with pytest.raises(ValueError):
        sep.inference(x, e[:3])
