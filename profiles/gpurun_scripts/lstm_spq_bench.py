"""Step time of lstm_tc_kernel against the sequences-per-gate-warp choice (PS_LSTM_SPQ, read once per process):
one full wave of 148 CTAs, L = 100 steps.  usage: lstm_spq_bench.py H"""
import os, sys
import torch
sys.path.insert(0, ".")
from puresound_b200 import ops
ops.require_device()
H = int(sys.argv[1]); spq = int(os.environ["PS_LSTM_SPQ"]); D = 1; L = 100
n_seq = 148 * 8 * spq
g = torch.Generator().manual_seed(0)
w_hh_t = (0.15 * (2 * torch.rand(D, H, 4 * H, generator=g) - 1)).cuda()
pk = ops.lstm_pack_weights(w_hh_t, H, D)
gx = (2 * torch.rand(n_seq * L, D * 4 * H, generator=g) - 1).cuda()
kw = dict(n_seq=n_seq, L=L, H=H, D=D, inner=1, outer_stride=L, inner_stride=0, step_stride=1, w_packed=pk, gx_interleaved=True)
for _ in range(3): ops.lstm(gx, w_hh_t, **kw)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize(); ev[0].record()
for _ in range(10): ops.lstm(gx, w_hh_t, **kw)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
print(f"H={H} spq={spq} seqs/CTA={8*spq}: {ms*1e3:.1f} us per launch, {ms*1e3/L:.2f} us per step, {ms*1e6/L/(8*spq):.0f} ns per step and sequence")
