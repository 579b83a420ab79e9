"""Phase timeline of the persistent hop kernel (experiments build: PS_B200_LIB=...libpuresound_b200_exp.so): per-phase work
time and grid-barrier time of CTA 0 for S concurrent streams."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from puresound_b200 import _lib, ops, testing  # noqa: E402
from puresound_b200.nnet.base_nn import SoTaskWrapModule  # noqa: E402
from puresound_b200.nnet.lobe.encoder import FreeEncDec  # noqa: E402
from puresound_b200.streaming.conv_tasnet_inference import StreamingConvTasNet, StreamingSeparator  # noqa: E402

ops.require_device()
torch.manual_seed(0)
m = SoTaskWrapModule(FreeEncDec(320, 512, 160), StreamingConvTasNet(512, 0, False, tcn_dim=512, per_tcn_stack=8, repeat_tcn=3, tcn_with_embed=[0] * 8,
                     tcn_norm="cLN", dconv_norm="cLN", causal=True), mask_constraint="ReLU", verbose=False).eval().cuda()
lib = _lib.load()
for S in (1, 16, 256):
    sep = StreamingSeparator(m, use_graph=False)
    sep.init_status(S)
    x = testing.white(S, 160, amp=0.1, seed=1).cuda()
    for _ in range(6):
        sep.step_wave(x)
    torch.cuda.synchronize()
    n = 2 + 2 * 99
    buf = (C.c_ulonglong * n)()
    lib.ps_debug_hop_times.argtypes = [C.c_void_p, C.c_int]
    rc = lib.ps_debug_hop_times(buf, n)
    t = [buf[i] for i in range(n)]
    work = [t[i + 1] - t[i] for i in range(0, n - 1, 2)]     # start->before barrier k, after barrier k -> before barrier k+1
    bar = [t[i + 1] - t[i] for i in range(1, n - 1, 2)]
    names = ["push", "encoder"] + [p for _ in range(24) for p in ("A", "B", "C", "E")] + ["decoder"]
    tot = (t[-1] - t[0]) / 1e3
    print(f"S={S}: CTA 0 timeline {tot:.1f} us over {len(bar)} barriers; mean barrier {sum(bar)/len(bar)/1e3:.2f} us")
    for ph in ("A", "B", "C", "E"):
        w = [work[i] for i in range(len(names)) if names[i] == ph]
        print(f"   phase {ph}: mean work {sum(w)/len(w)/1e3:.2f} us (min {min(w)/1e3:.2f}, max {max(w)/1e3:.2f})")
    print("   push %.2f enc %.2f dec %.2f us; first 12 barriers:" % (work[0] / 1e3, work[1] / 1e3, work[-1] / 1e3), [round(b / 1e3, 2) for b in bar[:12]])
