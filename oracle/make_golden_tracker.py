"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/tracker_seq.json by running the UNMODIFIED reference tracker
(`/root/reference/src/lib/utils/tracker.py`, through `oracle/ref_shims.py`) on a seeded synthetic detection sequence.

    python -m oracle.make_golden_tracker            # needs /root/reference; rewrites the fixture

`make_sequence()` is importable without the reference: the tests rebuild the same inputs from the seed and feed them to
`oracle/tracker_ref.py`.  The per-frame flow mirrors `detectors/base_detector.py:495-664` for `tracking_task`:
merge_outputs result -> gaussian_fusion (a closure inside run(), not importable: both arms use the restated formula)
-> first PnP per detection (`pnp_shell`, rep_mode 1 point assembly :558-566) -> `Tracker.step(results, boxes)`.
"""
import copy
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "tracker_seq.json")

K_CAM = np.array([[615.0, 0.0, 256.0], [0.0, 615.0, 256.0], [0.0, 0.0, 1.0]])
WIDTH = HEIGHT = 512


def _rot(rv):
    th = np.linalg.norm(rv)
    if th < 1e-12:
        return np.eye(3)
    k = rv / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def _vertices(scale):
    s = np.asarray(scale, np.float64) / scale[1]
    w, h, d = s / 2
    return np.array([[-w, -h, -d], [-w, -h, d], [-w, h, -d], [-w, h, d], [w, -h, -d], [w, -h, d], [w, h, -d], [w, h, d]])


def make_sequence(seed=7, n_frames=7):
    """-> (meta, frames); frames[f] = list of detection dicts as they look after merge_outputs (base_detector.py:495)."""
    rng = np.random.default_rng(seed)
    meta = {"camera_matrix": K_CAM.copy(), "width": WIDTH, "height": HEIGHT}
    objs = []
    for o in range(4):
        objs.append({
            "scale": np.array([rng.uniform(0.5, 1.2), 1.0, rng.uniform(0.5, 1.2)]) * rng.uniform(0.2, 0.4),
            "rv": rng.normal(size=3) * 0.5, "drv": rng.normal(size=3) * 0.02,
            "t": np.array([rng.uniform(-0.6, 0.6), rng.uniform(-0.4, 0.4), rng.uniform(2.0, 3.5)]),
            "dt": np.array([rng.uniform(-0.03, 0.03), rng.uniform(-0.02, 0.02), rng.uniform(-0.03, 0.03)]),
            "first": 0 if o < 3 else 2,                    # object 3 enters at frame 2
            "gone": (3, 4) if o == 2 else (),              # object 2 is missed in frames 3-4 and comes back
        })
    frames = []
    prev_kps = {}
    prev_ct = {}
    for f in range(n_frames):
        dets = []
        for oi, ob in enumerate(objs):
            if f < ob["first"] or f in ob["gone"]:
                continue
            R = _rot(ob["rv"] + f * ob["drv"])
            t = ob["t"] + f * ob["dt"]
            P = _vertices(ob["scale"]) @ R.T + t
            uv = (K_CAM @ P.T).T
            kps = uv[:, :2] / uv[:, 2:]
            noisy = lambda s: (kps + rng.normal(size=kps.shape) * s).reshape(-1)      # noqa: E731
            lo, hi = kps.min(0), kps.max(0)
            pad = 0.1 * (hi - lo)
            bbox = [float(lo[0] - pad[0]), float(lo[1] - pad[1]), float(hi[0] + pad[0]), float(hi[1] + pad[1])]
            ct = [(bbox[0] + bbox[2]) / 2, (bbox[1] + bbox[3]) / 2]
            tracking = (np.array(prev_ct[oi]) - np.array(ct) + rng.normal(size=2) * 0.3) if oi in prev_ct else np.zeros(2)
            tracking_hp = ((prev_kps[oi] - kps).reshape(-1) + rng.normal(size=16) * 0.3) if oi in prev_kps else np.zeros(16)
            hm_mean = noisy(0.8)
            hm_std = rng.uniform(0.5, 2.0, size=16)
            miss = rng.random(8) < 0.15                      # heat-map peak not found for this keypoint
            for j in np.nonzero(miss)[0]:
                hm_mean[2 * j:2 * j + 2] = -10000.0
                hm_std[2 * j:2 * j + 2] = -10000.0 * 0.32
            det = {
                "score": float(rng.uniform(0.45, 0.95)), "cls": 0, "bbox": bbox, "ct": ct,
                "tracking": tracking.astype(np.float64), "tracking_hp": tracking_hp.astype(np.float64),
                "kps": noisy(0.5), "kps_displacement_mean": noisy(1.0),
                "kps_displacement_std": rng.uniform(1.0, 3.0, size=16),
                "kps_heatmap_mean": hm_mean, "kps_heatmap_std": hm_std,
                "kps_heatmap_height": rng.uniform(0.3, 0.9, size=8),
                "obj_scale": (ob["scale"] / ob["scale"][1] * (1 + rng.normal(size=3) * 0.03)).astype(np.float32),
                "obj_scale_uncertainty": rng.uniform(0.05, 0.2, size=3).astype(np.float32),
            }
            dets.append(det)
            prev_kps[oi] = kps
            prev_ct[oi] = ct
        if f == 5:                                          # a weak detection that must not start a track
            d = copy.deepcopy(dets[0])
            d["score"] = 0.05
            d["ct"] = [d["ct"][0] + 150.0, d["ct"][1] + 120.0]
            d["bbox"] = [b + (150.0 if i % 2 == 0 else 120.0) for i, b in enumerate(d["bbox"])]
            dets.append(d)
        frames.append(dets)
    return meta, frames


def assemble_points(det):
    """rep_mode 1 (base_detector.py:558-566): interleave displacement and heat-map estimates."""
    pts = []
    dm = np.asarray(det["kps_displacement_mean"]).reshape(-1, 2)
    hm = np.asarray(det["kps_heatmap_mean"]).reshape(-1, 2)
    for i in range(8):
        pts.append(dm[i])
        pts.append(hm[i])
    return np.array(pts)


def summarize(ret, boxes):
    out = {"n_boxes": len(boxes), "tracks": []}
    for tr in ret:
        e = {"tracking_id": int(tr["tracking_id"]), "age": int(tr["age"]), "active": int(tr["active"]),
             "score": float(tr["score"])}
        for k in ("kps_mean_kf", "kps_std_kf", "obj_scale_kf", "obj_scale_uncertainty_kf", "location", "quaternion_xyzw",
                  "kps_pnp_kf"):
            if k in tr:
                e[k] = np.asarray(tr[k], np.float64).reshape(-1).tolist()
        out["tracks"].append(e)
    return out


def run_reference():
    sys.path.insert(0, ROOT)
    from oracle import ref_shims, tracker_ref
    ref_shims.install()
    opt = ref_shims.make_opt("dla_34", tracking_task=True, rep_mode=1, c="chair")
    from lib.utils.tracker import Tracker
    from lib.utils.pnp.cuboid_pnp_shell import pnp_shell
    meta, frames = make_sequence()
    trk = Tracker(opt)
    trk.init_track(meta)
    golden = []
    for dets in frames:
        results = copy.deepcopy(dets)
        boxes = []
        for det in results:
            m, s = tracker_ref.gaussian_fusion(det, opt.hps_uncertainty)
            det["kps_fusion_mean"], det["kps_fusion_std"] = m, s
            r = pnp_shell(opt, meta, det, assemble_points(det), det["obj_scale"], OPENCV_RETURN=opt.show_axes)
            if r is not None:
                boxes.append(r)
        ret, bx = trk.step(results, boxes)
        golden.append(summarize(ret, bx))
    opts = {k: getattr(opt, k) for k in ("kalman", "scale_pool", "hungarian", "use_pnp", "new_thresh", "max_age", "R", "c",
                                         "show_axes", "hps_uncertainty")}
    opts["conf_border"] = opt.conf_border[opt.c]
    return {"opt": opts, "frames": golden}


RENDER_OUT = os.path.join(ROOT, "tests", "golden", "track_render.npz")
RENDER_CASES = [          # (frames stepped before rendering, network input h, w)
    ("f2_512", 2, 512, 512),
    ("f4_256", 4, 256, 256),       # object 2 is lost here: lost tracks are rendered too
    ("f6_320x384", 6, 320, 384),
]


def render_meta(inp_h, inp_w):
    """meta as pre_process (fix_res) builds it for the 512 x 512 frames of the sequence and a network input of
    inp_h x inp_w: c = image centre, s = max side; trans_* through the reference's own get_affine_transform."""
    from lib.utils.image import get_affine_transform
    c = np.array([WIDTH / 2., HEIGHT / 2.], dtype=np.float32)
    s = max(HEIGHT, WIDTH) * 1.0
    return {"c": c, "s": s, "height": HEIGHT, "width": WIDTH, "inp_height": inp_h, "inp_width": inp_w,
            "out_height": inp_h // 4, "out_width": inp_w // 4, "camera_matrix": K_CAM.copy(),
            "trans_input": get_affine_transform(c, s, 0, [inp_w, inp_h]),
            "trans_output": get_affine_transform(c, s, 0, [inp_w // 4, inp_h // 4])}


def run_reference_render():
    """The UNMODIFIED `BaseDetector._get_additional_inputs` (base_detector.py:150-388) on the reference tracker's own
    track dicts after N frames of the seeded sequence -> previous-frame heat maps."""
    import types
    import torch
    sys.path.insert(0, ROOT)
    from oracle import ref_shims, tracker_ref
    ref_shims.install()
    opt = ref_shims.make_opt("dla_34", tracking_task=True, rep_mode=1, c="chair")
    opt.device = torch.device("cpu")
    from lib.utils.tracker import Tracker
    from lib.utils.pnp.cuboid_pnp_shell import pnp_shell
    from lib.detectors.base_detector import BaseDetector
    stub = types.SimpleNamespace(opt=opt)
    stub._trans_bbox = lambda *a: BaseDetector._trans_bbox(stub, *a)
    out = {}
    for name, nframes, ih, iw in RENDER_CASES:
        meta, frames = make_sequence()
        trk = Tracker(opt)
        trk.init_track(meta)
        for dets in frames[:nframes]:
            results = copy.deepcopy(dets)
            boxes = []
            for det in results:
                m, sd = tracker_ref.gaussian_fusion(det, opt.hps_uncertainty)
                det["kps_fusion_mean"], det["kps_fusion_std"] = m, sd
                r = pnp_shell(opt, meta, det, assemble_points(det), det["obj_scale"], OPENCV_RETURN=opt.show_axes)
                if r is not None:
                    boxes.append(r)
            trk.step(results, boxes)
        rm = render_meta(ih, iw)
        hm, hm_hp, _ = BaseDetector._get_additional_inputs(stub, trk.tracks, rm, with_hm=True, with_hm_hp=True)
        # fixture size: the smallest case is stored whole; the larger ones as every 4th pixel + per-channel sum / non-zero
        # count (a wrong radius, centre or heat changes all of them)
        for key, v in (("_hm", hm.numpy()[0]), ("_hm_hp", hm_hp.numpy()[0])):
            if ih * iw <= 256 * 256:
                out[name + key] = v
            else:
                out[name + key + "_sub4"] = np.ascontiguousarray(v[:, ::4, ::4])
                out[name + key + "_sum"] = v.astype(np.float64).sum(axis=(1, 2))
                out[name + key + "_nnz"] = (v != 0).sum(axis=(1, 2)).astype(np.int64)
        out[name + "_trans_input"] = rm["trans_input"]
        print(name, "tracks", len(trk.tracks), "hm max", float(hm.max()), "hm_hp max per joint",
              [round(float(v), 3) for v in hm_hp[0].amax(dim=(1, 2))])
    out["pre_thresh"] = float(opt.pre_thresh)
    out["render_hm_mode"] = int(opt.render_hm_mode)
    out["render_hmhp_mode"] = int(opt.render_hmhp_mode)
    return out


if __name__ == "__main__":
    g = run_reference()
    with open(OUT, "w") as f:
        json.dump(g, f)
    print("wrote", OUT, "frames", len(g["frames"]), [len(fr["tracks"]) for fr in g["frames"]])
    r = run_reference_render()
    np.savez_compressed(RENDER_OUT, **r)
    print("wrote", RENDER_OUT)
