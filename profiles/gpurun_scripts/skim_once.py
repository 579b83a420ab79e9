"""One eager forward of a SkiM recipe at 32 x (10 s + 6 s) (for ncu captures; PS_CUDA_GRAPH=0)."""
import sys
import torch
sys.path.insert(0, ".")
import bench
w = sys.argv[1] if len(sys.argv) > 1 else "tse_skim_v0_causal"
m = bench.build_model(w).to("cuda")
m.use_cuda_graph = False
mix, enr = bench.build_inputs(w, 0)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    y = m.inference(mix.cuda(), enr.cuda())
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
