"""Decode + grouping + soft-NMS + PnP on planted heads (4 objects per frame), for ncu / timing:
   python scripts/profile_decode.py [batch] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import centerpose_b200 as cpb  # noqa: E402
from centerpose_b200 import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
hb, truths = synth.planted_batch(B, n_obj=4, seed=5, disagree_px=1.0)
prm = cpb.decode_params(cpb.default_opt("dla_34"))
meta = cpb.make_meta(B, np.array([256., 256.], np.float32), 512.0, 512, 512, truths[0]["cam"]).cuda()
heads = {k: torch.from_numpy(v).cuda() for k, v in hb.items()}
for _ in range(3):
    _, poses, n_valid = cpb.decode_pnp(heads, meta, prm, want_dets=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    _, poses, n_valid = cpb.decode_pnp(heads, meta, prm, want_dets=False)
e1.record()
torch.cuda.synchronize()
print("batch %d: %.1f us per call, %.1f us per frame, detections %s" % (B, 1e3 * e0.elapsed_time(e1) / reps,
      1e3 * e0.elapsed_time(e1) / reps / B, n_valid.cpu().numpy()[:8]))
