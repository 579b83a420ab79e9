"""Per-kernel device time of ONE eager forward (torch profiler; no ncu). usage: model_breakdown.py workload"""
import sys
import torch
sys.path.insert(0, ".")
import bench
from torch.profiler import profile, ProfilerActivity
w = sys.argv[1]
m = bench.build_model(w).to("cuda")
m.use_cuda_graph = False
mix, enr = bench.build_inputs(w, 0)
mix = mix.cuda(); enr = None if enr is None else enr.cuda()
for _ in range(2): y = m.inference(mix, enr)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    y = m.inference(mix, enr); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=70))
