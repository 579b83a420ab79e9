"""ShardedSeparator: one host batch split over several devices, results gathered in order and identical
to the single-device forward.  Runs with two replicas on two streams of ONE device everywhere, and across two GPUs
where the box has them."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(name="cfg1"):
    from puresound_b200 import ops, recipes, testing

    ops.require_device()
    torch.manual_seed(0)
    m = recipes.baseline_config(name).eval()
    testing.perturb_(m, seed=1)
    return m.to("cuda:0")


@pytest.mark.parametrize("devices", [[0, 0], [0, 1]])
def test_sharded_equals_single_device(devices):
    from puresound_b200 import testing
    from puresound_b200.sharding import ShardedSeparator

    if max(devices) >= torch.cuda.device_count():
        pytest.skip("needs two GPUs")
    m = _model()
    sep = ShardedSeparator(m, devices)
    for n in (5, 1, 4):  # uneven split, fewer items than devices, even split
        x = testing.noisy_speech(n, 16000, seed=100 + n)[0]
        want = m.inference(x)
        for rep in range(3):  # eager, capture, replay on every replica
            got = sep.inference(x)
            assert got.shape == want.shape and got.is_pinned()
            assert torch.equal(got, want), f"n={n} call {rep}"


def test_sharded_tse_pairs():
    from puresound_b200 import recipes, testing
    from puresound_b200.sharding import ShardedSeparator

    torch.manual_seed(0)
    m = recipes.init_model("td_tse_conv_tasnet_v0", verbose=False).eval().to("cuda:0")
    sep = ShardedSeparator(m, [0, min(1, torch.cuda.device_count() - 1)])
    x, e = testing.noisy_speech(3, 16000, seed=5)[0], testing.noisy_speech(3, 24000, seed=6)[0]
    want = m.inference(x, e)
    for _ in range(3):
        assert torch.equal(sep.inference(x, e, reuse_output=True), want)
