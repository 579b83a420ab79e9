"""Encoder (FreeEncDec analysis, win 32 / hop 16) at the cfg2 and cfg3 shapes: tcgen05 K = 32 path vs the exact-fp32
register filterbank kernel.  CUDA events, 20 launches each after 5 warm-ups; prints ms and achieved output GB/s."""
import sys
import torch
sys.path.insert(0, ".")
from puresound_b200 import ops
ops.require_device()
for N, L, M in ((64, 64000, 512), (32, 160000, 128)):
    g = torch.Generator().manual_seed(0)
    wav = (0.1 * (2 * torch.rand(N, L, generator=g) - 1)).cuda()
    w = (0.2 * (2 * torch.rand(M, 32, generator=g) - 1)).cuda()
    T = (L - 32) // 16 + 1
    pk = ops.pack_weights(w, M, 32, 32)
    out = torch.empty(N, T, M, device="cuda")
    for name, kw in (("tcgen05", dict(w_packed=pk, backend=ops.GEMM_TCGEN05)), ("simt", dict(backend=ops.GEMM_SIMT))):
        f = lambda: ops.gemm(wav, w, batch=N, rows=T, M=M, K=32, x_batch_stride=L, x_row_stride=16, w_row_stride=32, out=out, **kw)
        for _ in range(5):
            f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"encoder {N}x{T}x{M}x32 {name}: {ms:.4f} ms, {N * T * M * 4 / ms / 1e6:.0f} GB/s written")
