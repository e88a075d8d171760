"""Per-op timing (cp_plan_profile) of one forward: python scripts/profile_ops.py [batch] [precision] [tracking 0/1]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import centerpose_b200 as cpb  # noqa: E402
from centerpose_b200 import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
trk = len(sys.argv) > 3 and sys.argv[3] == "1"
opt = cpb.default_opt("dla_34", tracking_task=True) if trk else cpb.default_opt("dla_34")
m = cpb.create_model(opt.arch, opt.heads, opt.head_conv, opt)
m.precision = prec
m.load_state_dict(synth.seeded_state_dict(m, seed=0, offset_std=0.3))
m = m.cuda().eval()
x = torch.from_numpy(synth.normalize_frames(synth.synthetic_frames(B, 512, 512, seed=1))).cuda()
eng = m.engine(B, 512, 512, x.device)
args = (x,)
if trk:
    args = (x, x.clone(), torch.zeros(B, 1, 512, 512, device=x.device), torch.zeros(B, 8, 512, 512, device=x.device))
for _ in range(3):
    ops = eng.profile(*args)
reps = [eng.profile(*args) for _ in range(8)]
ms = np.mean([[o["ms"] for o in r] for r in reps], axis=0)
tot = ms.sum()
print("batch %d, %s%s: sum of ops %.3f ms (%d ops)" % (B, prec, ", tracking" if trk else "", tot, len(ops)))
groups = {}
for o, t in zip(ops, ms):
    n = o["name"]
    if "offset_mask" in n: g = "offset convs"
    elif "(dcn)" in n: g = "dcn"
    elif n.startswith(("base.base_layer", "base.level0", "base.level1", "base.pre_")): g = "stem+l0+l1"
    elif n.startswith("base."): g = "base " + n.split(".")[1]
    elif "merged" in n or n.split(".")[0] in opt.heads: g = "heads"
    elif ".up_" in n: g = "upsample"
    else: g = n
    groups[g] = groups.get(g, 0) + t
for g, v in sorted(groups.items(), key=lambda kv: -kv[1])[:12]:
    print("  %-16s %7.3f ms %5.1f %%" % (g, v, 100 * v / tot))
if os.environ.get("CP_PROFILE_TOP"):
    for o, t in sorted(zip(ops, ms), key=lambda kv: -kv[1])[:int(os.environ["CP_PROFILE_TOP"])]:
        print("  %-50s %7.3f ms" % (o["name"], t))
