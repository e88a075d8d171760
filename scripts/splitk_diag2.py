"""Reproduces the call order of tests/test_gpu_bench_parity.py::test_512_frame17_of_b32."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import centerpose_b200 as cpb
from centerpose_b200 import synth

opt = cpb.default_opt("dla_34")
m = cpb.create_model(opt.arch, opt.heads, opt.head_conv, opt)
m.precision = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
m.load_state_dict(synth.seeded_state_dict(m, seed=12, offset_std=0.3, head_gain=1.0))
m = m.cuda().eval()
frames = synth.synthetic_frames(32, 512, 512, seed=4242)
x = torch.from_numpy(synth.normalize_frames(frames))
full = m(x.cuda())[-1]
again = m(x.cuda())[-1]
one = m(x[17:18].contiguous().cuda())[-1]
one2 = m(x[17:18].contiguous().cuda())[-1]
print("aliasing full/again:", full["hm"].data_ptr() == again["hm"].data_ptr(), " one/one2:", one["hm"].data_ptr() == one2["hm"].data_ptr())
for h in opt.heads:
    d = (full[h][17:18] - one[h]).abs().max().item() / one[h].abs().max().item()
    d2 = (full[h][17:18] - one2[h]).abs().max().item() / one[h].abs().max().item()
    print(h, "b32[17] vs alone %.2e  (second call %.2e)" % (d, d2))
