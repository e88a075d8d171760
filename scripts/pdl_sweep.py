"""Forward time vs batch with PDL forced on / off (CP_PDL=1 / CP_NO_PDL=1): python scripts/pdl_sweep.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, torch
sys.path.insert(0, %r)
import centerpose_b200 as cpb
from centerpose_b200 import synth
opt = cpb.default_opt("dla_34")
m = cpb.create_model(opt.arch, opt.heads, opt.head_conv, opt)
m.load_state_dict(synth.seeded_state_dict(m, seed=0, offset_std=0.3))
m = m.cuda().eval()
res = []
for B in (1, 2, 4, 8, 16):
    x = torch.from_numpy(synth.normalize_frames(synth.synthetic_frames(B, 512, 512, seed=1))).cuda()
    eng = m.engine(B, 512, 512, x.device)
    for _ in range(5): eng.forward(x)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): eng.forward(x)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    res.append("b%%d %%.3f" %% (B, best))
print(" ".join(res))
''' % ROOT
for name, env in (("pdl on ", {"CP_PDL": "1"}), ("pdl off", {"CP_NO_PDL": "1"})):
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, "-c", CODE], env=e, capture_output=True, text=True).stdout.strip().splitlines()
    print(name, out[-1] if out else "?")
