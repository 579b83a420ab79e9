"""CUDA-graph replay behind SoTaskWrapModule.inference(): same results as eager launches for changing inputs, host and
device tensors, and invalidation when a parameter changes (load_state_dict / in-place update)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    from puresound_b200 import ops, recipes, testing

    ops.require_device()
    torch.manual_seed(0)
    m = recipes.baseline_config("cfg1").eval()
    testing.perturb_(m, seed=1)
    return m.to("cuda")


def _eager(model, x):
    g, model.use_cuda_graph = model.use_cuda_graph, False
    try:
        return model.inference(x)
    finally:
        model.use_cuda_graph = g


def test_graph_replay_matches_eager_for_new_inputs(model):
    from puresound_b200 import ops, testing

    model.use_cuda_graph = True
    model._graphs.clear()
    xs = [testing.noisy_speech(2, 16000, seed=s)[0] for s in range(5)]
    launches = []
    for i, x in enumerate(xs):
        n0 = ops.launch_count
        y = model.inference(x.cuda() if i % 2 else x)  # host and device inputs alternate
        launches.append(ops.launch_count - n0)
        assert y.is_cuda == bool(i % 2)
        assert torch.equal(y.cpu(), _eager(model, x).cpu()), f"call {i}"
    # call 0 runs eagerly, call 1 captures (its kernels are recorded, not counted twice), calls 2.. only replay
    assert launches[0] > 100 and launches[2] == 0 and launches[4] == 0


def test_graph_is_dropped_when_a_parameter_changes(model):
    from puresound_b200 import testing

    model.use_cuda_graph = True
    x = testing.noisy_speech(1, 16000, seed=11)[0].cuda()
    for _ in range(3):
        y0 = model.inference(x)
    p = model.masker.tcn_list[0][0].out_conv.weight
    with torch.no_grad():
        p.mul_(1.5)
    try:
        y1 = model.inference(x)
        assert not torch.equal(y0, y1)
        assert torch.equal(y1, _eager(model, x))
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        for _ in range(3):
            model.inference(x)  # captured again with the new weights
        model.load_state_dict(sd)  # copy_ into the same storage bumps the version counters
        assert torch.equal(model.inference(x), y1)
    finally:
        with torch.no_grad():
            p.div_(1.5)


def test_input_shape_slots_are_bounded(model):
    from puresound_b200 import testing

    model.use_cuda_graph = True
    for L in (8000, 9000, 10000, 11000):
        x = testing.noisy_speech(1, L, seed=L)[0].cuda()
        for _ in range(3):
            y = model.inference(x)
        assert torch.equal(y, _eager(model, x))
    assert len(model._graphs) <= model._GRAPH_SLOTS


def test_inference_stream_equals_inference(model):
    """The serving loop (H2D / forward / D2H on three streams) returns, in order, exactly what inference() returns for each
    host batch - pinned and pageable inputs, a generator as the source, more batches than the pipeline depth."""
    from puresound_b200 import recipes, testing

    model.use_cuda_graph = True
    xs = [testing.noisy_speech(2, 16000, seed=20 + s)[0] for s in range(7)]
    xs = [x.pin_memory() if i % 2 else x for i, x in enumerate(xs)]
    ys = list(model.inference_stream(x for x in xs))
    assert len(ys) == len(xs)
    for i, (x, y) in enumerate(zip(xs, ys)):
        assert not y.is_cuda and y.is_pinned()
        assert torch.equal(y, model.inference(x)), f"batch {i}"
    assert list(model.inference_stream([])) == []
    # results from the ring of pinned buffers: each one is valid when it is yielded
    for i, (x, y) in enumerate(zip(xs, model.inference_stream(xs, depth=1, reuse_host_buffers=True))):
        assert torch.equal(y, ys[i]), f"ring batch {i}"
    # batches of changing shape in one stream (the staging slots are re-made, the graph cache holds two shapes)
    mixed = [testing.noisy_speech(2 if i % 3 else 1, 16000 if i % 2 else 24000, seed=70 + i)[0] for i in range(6)]
    for i, (x, y) in enumerate(zip(mixed, model.inference_stream(mixed))):
        assert torch.equal(y, model.inference(x)), f"mixed batch {i}"
    # (noisy, enroll) pairs through a TSE model
    torch.manual_seed(0)
    tse = recipes.init_model("td_tse_conv_tasnet_v0", verbose=False).eval().to("cuda")
    pairs = [(testing.noisy_speech(1, 16000, seed=40 + s)[0], testing.noisy_speech(1, 24000, seed=50 + s)[0]) for s in range(4)]
    for (x, e), y in zip(pairs, tse.inference_stream(pairs, depth=1)):
        assert torch.equal(y, tse.inference(x, e))


def test_stft_model_alternating_lengths_replays_the_right_window_table():
    """ConvEncDec caches the iSTFT window-sum-square table per frame count and a captured graph reads it by raw pointer:
    two input lengths alternating (both graphs alive) must each keep their own table (the cache used to be
    replaced, freeing the table under the other graph)."""
    from puresound_b200 import recipes, testing

    torch.manual_seed(0)
    m = recipes.baseline_config("cfg4").eval()
    testing.perturb_(m, seed=1)
    m = m.to("cuda")
    m.use_cuda_graph = True
    enr = testing.noisy_speech(1, 24000, seed=3)[0].cuda()
    xa = [testing.noisy_speech(1, 16000, seed=60 + i)[0].cuda() for i in range(4)]
    xb = [testing.noisy_speech(1, 20480, seed=80 + i)[0].cuda() for i in range(4)]
    for i in range(4):  # A eager, B eager, A capture, B capture, A replay, B replay, ...
        for x in (xa[i], xb[i]):
            y = m.inference(x, enr)
            junk = torch.full((1, y.shape[1]), 7.0, device="cuda")  # same size class as a freed table would be
            g, m.use_cuda_graph = m.use_cuda_graph, False
            want = m.inference(x, enr)
            m.use_cuda_graph = g
            assert torch.equal(y, want), f"round {i}, L={x.shape[1]}"
            del junk
    assert len(m._graphs) == 2


def test_mode_or_constraint_change_is_not_replayed(model):
    """A captured graph is keyed on the train/eval flags and the mask / output constraints: changing them after capture
    must not replay the old forward."""
    from puresound_b200 import testing

    model.use_cuda_graph = True
    model.eval()
    x = testing.noisy_speech(1, 16000, seed=12)[0].cuda()
    for _ in range(3):
        y0 = model.inference(x)
    old = model.mask_constraint
    try:
        model.mask_constraint = "sigmoid"
        y1 = model.inference(x)
        assert not torch.equal(y0, y1)
        assert torch.equal(y1, _eager(model, x))
    finally:
        model.mask_constraint = old
    assert torch.equal(model.inference(x), y0)
