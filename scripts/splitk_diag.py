"""Split-K diagnosis: the same frame at batch 1 with split-K on / off (conv_tma, dcn_tma separately), run twice."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import centerpose_b200 as cpb
from centerpose_b200 import synth

opt = cpb.default_opt("dla_34")
m = cpb.create_model(opt.arch, opt.heads, opt.head_conv, opt)
m.precision = "tf32x3"
m.load_state_dict(synth.seeded_state_dict(m, seed=12, offset_std=0.3, head_gain=1.0))
m = m.cuda().eval()
seed, idx = int(sys.argv[1]) if len(sys.argv) > 1 else 4242, int(sys.argv[2]) if len(sys.argv) > 2 else 17
frames = synth.synthetic_frames(32, 512, 512, seed=seed)
x = torch.from_numpy(synth.normalize_frames(frames)).cuda()
one = x[idx:idx + 1].contiguous()


def run(env, inp=one):
    if env is None:
        os.environ.pop("CP_NO_SPLITK", None)
    else:
        os.environ["CP_NO_SPLITK"] = env
    out = m(inp)[-1]
    torch.cuda.synchronize()
    return {h: out[h].clone() for h in opt.heads}


ref = run("1")
full = run("1", x)
for name, env in (("split conv+dcn", None), ("split conv+dcn again", None), ("split dcn only", "conv"), ("split conv only", "dcn"),
                  ("no split again", "1")):
    got = run(env)
    line = "%-22s" % name
    for h in opt.heads:
        d = (got[h] - ref[h]).abs().max().item() / ref[h].abs().max().item()
        line += " %s %.1e" % (h, d)
    print(line)
line = "%-22s" % "b32[idx] vs alone"
for h in opt.heads:
    d = (full[h][idx:idx + 1] - ref[h]).abs().max().item() / ref[h].abs().max().item()
    line += " %s %.1e" % (h, d)
print(line)
d = (full["hm"][idx:idx + 1] - run(None)["hm"]).abs()
print("hm: where", np.unravel_index(d.argmax().item(), d.shape), "max", d.max().item(), "count > 1e-3 of max:",
      int((d > 1e-3 * ref["hm"].abs().max()).sum().item()))
