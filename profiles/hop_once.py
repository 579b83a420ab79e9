"""A few eager hops of the persistent hop kernel at S streams (PROBE_S, default 256): the program `ncu -k regex:stream_hop` profiles."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from puresound_b200 import ops, testing  # noqa: E402
from puresound_b200.nnet.base_nn import SoTaskWrapModule  # noqa: E402
from puresound_b200.nnet.lobe.encoder import FreeEncDec  # noqa: E402
from puresound_b200.streaming.conv_tasnet_inference import StreamingConvTasNet, StreamingSeparator  # noqa: E402

ops.require_device()
torch.manual_seed(0)
S = int(os.environ.get("PROBE_S", 256))
m = SoTaskWrapModule(FreeEncDec(320, 512, 160), StreamingConvTasNet(512, 0, False, tcn_dim=512, per_tcn_stack=8, repeat_tcn=3, tcn_with_embed=[0] * 8,
                     tcn_norm="cLN", dconv_norm="cLN", causal=True), mask_constraint="ReLU", verbose=False).eval().cuda()
sep = StreamingSeparator(m, use_graph=False)
sep.init_status(S)
x = testing.white(S, 160, amp=0.1, seed=1).cuda()
for _ in range(8):
    y = sep.step_wave(x)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
