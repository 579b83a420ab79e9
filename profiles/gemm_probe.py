"""Time the three 1x1-conv GEMM variants of a cfg2 TCN block in isolation (back-to-back launches, CUDA events).
Used with the -DPS_EXPERIMENTS build (PS_B200_LIB=...libpuresound_b200_exp.so PS_WIDE_DBG=n) to see what each stream of
work costs; with the release build it is the per-kernel timing of run notes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from puresound_b200 import ops  # noqa: E402

ops.require_device()
B, T, C = int(os.environ.get("PROBE_B", 64)), int(os.environ.get("PROBE_T", 3999)), int(os.environ.get("PROBE_K", 512))
M = int(os.environ.get("PROBE_M", 512))
reps = int(os.environ.get("PROBE_REPS", 30))
g = torch.Generator().manual_seed(0)
x = (torch.rand(B, T, C, generator=g) - 0.5).cuda()
res = (torch.rand(B, T, M, generator=g) - 0.5).cuda()
w = (0.05 * (torch.rand(M, C, generator=g) - 0.5)).cuda()
bias = torch.rand(M, generator=g).cuda()
sc, sh = (torch.rand(B, C, generator=g) + 0.5).cuda(), (torch.rand(B, C, generator=g) - 0.5).cuda()
slope = torch.tensor([0.25]).cuda()
pk = ops.pack_weights(w, M, C, C)
pro = ops.Prologue(ops.PRO_AFFINE, ops.ACT_PRELU, sc, sh, C, None, slope)
out = torch.empty(B, T, M, device="cuda")
variants = {
    "in_conv  (no prologue, stats)": dict(want_stats=True),
    "pointwise(affine+PReLU, bias, stats)": dict(pro=pro, bias=bias, want_stats=True),
    "out_conv (affine+PReLU, bias, residual)": dict(pro=pro, bias=bias, residual=res),
}
ops.path_log = []
line = []
for name, kw in variants.items():
    for _ in range(3):
        ops.linear(x, w, w_packed=pk, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.linear(x, w, w_packed=pk, out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    line.append(f"{name}: {ms:.4f} ms")
print(f"dbg={os.environ.get('PS_WIDE_DBG', '-')} from2={os.environ.get('PS_PAIR_FROM2', '-')} B={B} T={T} M={M} K={C} path={ops.path_log[0][1]} | " + " | ".join(line))
