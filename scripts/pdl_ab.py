"""A/B of programmatic dependent launch (CP_NO_PDL=1 turns it off): forward-only and image -> pose latency at batch 1
(plain launches and CUDA graph) and batch 32.   python scripts/pdl_ab.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import centerpose_b200 as cpb  # noqa: E402
from centerpose_b200 import synth  # noqa: E402


def timeit(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


opt = cpb.default_opt("dla_34")
m = cpb.create_model(opt.arch, opt.heads, opt.head_conv, opt)
m.precision = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
m.load_state_dict(synth.seeded_state_dict(m, seed=0, offset_std=0.3))
m = m.cuda().eval()
prm = cpb.decode_params(opt)
cam = synth.default_camera(512, 512)
out = {}
for B in (1, 32):
    x = torch.from_numpy(synth.normalize_frames(synth.synthetic_frames(B, 512, 512, seed=1))).cuda()
    eng = m.engine(B, 512, 512, x.device)
    meta = cpb.make_meta(B, np.array([256., 256.], np.float32), 512.0, 512, 512, cam).cuda()
    heads = eng.forward(x)
    ref = {k: v.clone() for k, v in heads.items()}
    out["fwd_b%d" % B] = timeit(lambda: eng.forward(x), n=30 if B == 1 else 10)
    out["infer_b%d" % B] = timeit(lambda: eng.infer(x, meta, prm), n=30 if B == 1 else 10)
    if B == 1:
        g = cpb.InferGraph(eng, 1, prm)
        out["graph_b1"] = timeit(lambda: g(x, meta), n=50)
    h2 = eng.forward(x)
    out["same_b%d" % B] = all(torch.equal(ref[k], h2[k]) for k in ref)
print("PDL", "off" if os.environ.get("CP_NO_PDL") else "on", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items()})
