"""BASELINE config 5 leg alone (8 video streams, tracking dla_34 + device tracker): python scripts/track_b8.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import centerpose_b200 as cpb  # noqa: E402
from centerpose_b200 import synth  # noqa: E402

dev = torch.device("cuda")
cam = synth.default_camera(512, 512)
topt = cpb.default_opt("dla_34", tracking_task=True)
tm = cpb.create_model(topt.arch, topt.heads, topt.head_conv, topt)
tm.load_state_dict(synth.seeded_state_dict(tm, seed=0, offset_std=0.3, head_gain=1.0))
tdet = cpb.ObjectPoseDetector(topt, model=tm)
vids = [torch.from_numpy(synth.synthetic_frames(8, 512, 512, seed=500 + i)).to(dev) for i in range(4)]
xcal = torch.from_numpy(synth.normalize_frames(synth.synthetic_frames(8, 512, 512, seed=500))).to(dev)
teng = tdet.model.engine(8, 512, 512, dev)
z1, z8 = torch.zeros((8, 1, 512, 512), device=dev), torch.zeros((8, 8, 512, 512), device=dev)
synth.calibrate_head_bias(tdet.model, teng.forward(xcal, xcal, z1, z8), 4)
for i in range(6):
    tdet.run_batch(vids[i % 4], cam, track=True, to_host=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(12):
    tdet.run_batch(vids[i % 4], cam, track=True, to_host=False)
e1.record()
torch.cuda.synchronize()
_, nt = tdet.run_batch(vids[0], cam, track=True)
fw = []
for _ in range(3):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        teng.forward(xcal, xcal, z1, z8)
    b.record()
    torch.cuda.synchronize()
    fw.append(a.elapsed_time(b) / 5)
print("env %s: %.2f ms per step, forward only %.2f ms, tracks %.2f" % (
    {k: v for k, v in os.environ.items() if k.startswith("CP_")}, e0.elapsed_time(e1) / 12, min(fw), float(np.mean(nt))))
