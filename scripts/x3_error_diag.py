"""Where does the tf32x3 plan lose accuracy at 512 x 512, and what does each remedy cost?

For a list of environment configurations: error of the heads against the fp64 evaluation of the same graph
(tests/golden/_cache/net_dla34_b1_512_truth64.npz when present, else computed on the host) at batch 1, and the forward
time at batch 32.  Diagnostics only (GPU box):   python scripts/x3_error_diag.py [out.json]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import centerpose_b200 as cpb  # noqa: E402
from centerpose_b200 import synth  # noqa: E402
from tests.util import golden, net_case_inputs  # noqa: E402

CONFIGS = [
    ("fp32", "fp32", {}),
    ("tf32x3 (default: group 1, fused heads)", "tf32x3", {}),
    ("tf32x3 group=2", "tf32x3", {"CP_X3_GROUP": "2"}),
    ("tf32x3 group=3 (round-1 grouping)", "tf32x3", {"CP_X3_GROUP": "3"}),
    ("tf32x3 heads not fused", "tf32x3", {"CP_NO_FUSE_HEADS": "1"}),
    ("tf32x3 no igemm_umma", "tf32x3", {"CP_NO_UMMA": "1"}),
    ("tf32", "tf32", {}),
]
KNOBS = ("CP_X3_GROUP", "CP_NO_UMMA", "CP_NO_DCN_TMA", "CP_NO_FUSE_HEADS")


def truth_512(g, x, opt, sd):
    cache = os.path.join(ROOT, "tests", "golden", "_cache", "net_dla34_b1_512_truth64.npz")
    if os.path.exists(cache):
        z = np.load(cache)
        return {h: z[h] for h in opt.heads}
    from oracle import net_ref
    sd64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}
    out = net_ref.forward(torch.from_numpy(x).double(), sd64, opt.heads, "dla_34")
    return {h: v.numpy() for h, v in out.items()}


def main():
    g = golden("net_dla34_b1_512")
    x, _ = net_case_inputs(g)
    opt = cpb.default_opt("dla_34")
    m0 = cpb.create_model(opt.arch, opt.heads, opt.head_conv, opt)
    sd = synth.seeded_state_dict(m0, seed=int(g["wseed"]), offset_std=float(g["offset_std"]))
    truth = truth_512(g, x, opt, sd)
    xb = torch.from_numpy(synth.normalize_frames(synth.synthetic_frames(32, 512, 512, seed=5))).cuda()
    x1 = torch.from_numpy(x).cuda()
    rows = []
    for name, prec, env in CONFIGS:
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        m = cpb.create_model(opt.arch, opt.heads, opt.head_conv, opt)
        m.precision = prec
        m.load_state_dict(sd)
        m = m.cuda().eval()
        out = m(x1)[-1]
        errs, refd = {}, {}
        for h in opt.heads:
            mag = np.abs(truth[h]).max()
            errs[h] = float(np.abs(out[h].cpu().numpy().astype(np.float64) - truth[h]).max() / mag)
            refd[h] = float(np.abs(g["head_" + h].astype(np.float64) - truth[h]).max() / mag)
        eng = m.engine(32, 512, 512, xb.device)
        for _ in range(2):
            eng.forward(xb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.forward(xb)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        ops = eng.profile(xb)
        heads_ms = sum(o["ms"] for o in ops if "merged" in o["name"] or o["name"].split(".")[0] in opt.heads)
        dcn_ms = sum(o["ms"] for o in ops if "(dcn)" in o["name"])
        om_ms = sum(o["ms"] for o in ops if "conv_offset_mask" in o["name"])
        row = {"config": name, "worst_vs_fp64": max(errs.values()), "worst_ratio_to_ref": max(errs[h] / refd[h] for h in errs),
               "errs": errs, "forward_ms_b32": ms, "heads_ms": heads_ms, "dcn_ms": dcn_ms, "offset_conv_ms": om_ms}
        rows.append(row)
        print("%-34s worst gpu-vs-fp64 %.2e (%.1fx the reference's own fp32 distance)  fwd %.2f ms  heads %.2f  dcn %.2f  om %.2f"
              % (name, row["worst_vs_fp64"], row["worst_ratio_to_ref"], ms, heads_ms, dcn_ms, om_ms), flush=True)
        del m, eng
    print("reference fp32 vs fp64:", {h: "%.2e" % v for h, v in refd.items()})
    if len(sys.argv) > 1:
        json.dump(rows, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("done in %.0f s" % (time.time() - t0))
