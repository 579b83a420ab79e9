"""One eager forward of a tse_unet_tcn recipe (for ncu launch lists; PS_CUDA_GRAPH=0). usage: unet_once.py name batch reps"""
import sys
import torch
sys.path.insert(0, ".")
from puresound_b200 import recipes, testing, ops
ops.require_device()
name, n, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
torch.manual_seed(0)
m = recipes.init_model(name, verbose=False).eval()
testing.perturb_(m, seed=1)
m = m.to("cuda")
mix = testing.noisy_speech(n, 64000, seed=1)[0].cuda()
enr = testing.noisy_speech(n, 96000, seed=2)[0].cuda()
for _ in range(reps):
    y = m.inference(mix, enr)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
