"""Three batch-1 image -> pose calls (cp_infer) for an ncu launch list: `ncu --metrics gpu__time_duration.sum ... python scripts/infer_b1.py`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import centerpose_b200 as cpb
from centerpose_b200 import synth

opt = cpb.default_opt("dla_34")
m = cpb.create_model(opt.arch, opt.heads, opt.head_conv, opt)
m.load_state_dict(synth.calibrate_head_bias(synth.seeded_state_dict(m, seed=12, offset_std=0.3, head_gain=1.0)) if hasattr(synth, "calibrate_head_bias") and False
                  else synth.seeded_state_dict(m, seed=12, offset_std=0.3, head_gain=1.0))
det = cpb.ObjectPoseDetector(opt, model=m)
frames = synth.synthetic_frames(1, 512, 512, seed=5)
cam = synth.default_camera(512, 512)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    poses, n_valid = det.run_batch(frames, cam)
torch.cuda.synchronize()
print("n_valid", n_valid)
