# round 2, run 2: debug cfg1b batch-1, rest of the GPU tests (no -x), MMA-rate microbenchmark
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r02_run2_dbg.log 2>&1
import torch, traceback
from puresound_b200 import recipes, testing, ops, _lib
ops.require_device()
torch.manual_seed(0)
m = recipes.baseline_config("cfg1b").eval(); testing.perturb_(m, seed=1); m = m.to("cuda")
lib = _lib.load()
for n in (1, 2, 64, 1):
    x = testing.noisy_speech(n, 64000, seed=5)[0]
    try:
        y = m.inference(x); print("batch", n, "ok", float(y.abs().mean()))
    except Exception as e:
        print("batch", n, "FAILED", repr(e), lib.ps_last_cuda_error())
        traceback.print_exc()
PY
tail -30 gpurun_out/r02_run2_dbg.log
./profiles/microbench/mma_rate > gpurun_out/r02_run2_mma_rate.txt 2>&1; cat gpurun_out/r02_run2_mma_rate.txt
python -m pytest tests -m gpu -q > gpurun_out/r02_run2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run2_pytest.log
tail -15 gpurun_out/r02_run2_pytest.log
