"""Per-batch wall time of the serving loop (inference_stream) against sequential inference() on host batches."""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
w = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
m = bench.build_model(w).to("cuda")
mix, enr = bench.build_inputs(w, 0)
mix = mix.pin_memory(); enr = None if enr is None else enr.pin_memory()
for _ in range(4):
    m.inference(mix, enr)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    m.inference(mix, enr)
torch.cuda.synchronize()
print(w, "sequential ms/batch", round(1e3 * (time.perf_counter() - t0) / 10, 2))
for depth, reuse in ((2, False), (2, True), (1, True), (4, True)):
    for _ in m.inference_stream([(mix, enr)] * 6, depth=depth, reuse_host_buffers=reuse):
        pass
    torch.cuda.synchronize()
    ts = [time.perf_counter()]
    for _ in m.inference_stream(((mix, enr) for _ in range(24)), depth=depth, reuse_host_buffers=reuse):
        ts.append(time.perf_counter())
    torch.cuda.synchronize()
    d = [round(1e3 * (b - a), 1) for a, b in zip(ts, ts[1:])]
    print(w, f"depth={depth} reuse={reuse}: ms/batch {1e3 * (ts[-1] - ts[0]) / 24:.2f}; per yield {d}")
