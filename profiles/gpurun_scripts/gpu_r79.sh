mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "thin or filterbank or gemm" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_full.py -q -k "unet" 2>&1 | tail -3
python - <<'PY' > gpurun_out/r79_unet.log 2>&1
import torch, time
from puresound_b200 import recipes, testing, ops
ops.require_device()
for name, n in (("tse_unet_tcn_v0", 64),):
    torch.manual_seed(0)
    m = recipes.init_model(name, verbose=False).eval()
    testing.perturb_(m, seed=1)
    m = m.to("cuda")
    mix = testing.noisy_speech(n, 64000, seed=1)[0].cuda()
    enr = testing.noisy_speech(n, 96000, seed=2)[0].cuda()
    for _ in range(4): y = m.inference(mix, enr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): y = m.inference(mix, enr)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    print(f"{name} {n} x (4 s mix + 6 s enroll): {ms:.2f} ms/step = {n*4/(ms/1e3):.0f} audio-s/s, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
    # per-kernel-class device time of one eager forward (torch profiler, no ncu)
    m.use_cuda_graph = False
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        y = m.inference(mix, enr); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
PY
tail -32 gpurun_out/r79_unet.log | cut -c1-170
