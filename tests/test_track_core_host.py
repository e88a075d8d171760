"""CPU test of centerpose_b200/csrc/track_core.h (the tracker logic the CUDA kernel runs, compiled for the host):
fed with fp32 pose records built from the seeded detection sequence, it must reproduce what the UNMODIFIED reference
tracker produced (tests/golden/tracker_seq.json) -- same track ids / ages / box counts, filter state, pooled scale and
second-PnP pose.  Inputs are rounded to fp32 (the record type), so values agree to ~1e-6 relative, not bit-exactly."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

from centerpose_b200 import _lib as L
from oracle import make_golden_tracker as mg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "tracker_seq.json")
VISIBLE = {"book": 6, "chair": 6, "cereal_box": 6, "camera": 3, "bottle": 3, "cup": 3, "bike": 0, "laptop": 0, "shoe": 0}


@pytest.fixture(scope="module")
def track_host():
    src = os.path.join(ROOT, "tests", "host", "track_core_host.cpp")
    out_dir = os.path.join(ROOT, "tests", "host", "_build")
    so = os.path.join(out_dir, "libtrack_core_host.so")
    hdrs = [os.path.join(ROOT, "centerpose_b200", "csrc", h) for h in ("track_core.h", "pose_core.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.trk_create.restype = ctypes.c_void_p
    lib.trk_create.argtypes = [ctypes.c_int] * 5 + [ctypes.c_double] * 4 + [ctypes.c_int] * 3
    lib.trk_destroy.argtypes = [ctypes.c_void_p]
    lib.trk_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_double,
                             ctypes.c_double, ctypes.c_void_p]
    lib.trk_gaussian_radius.restype = ctypes.c_double
    lib.trk_gaussian_radius.argtypes = [ctypes.c_double, ctypes.c_double]
    lib.trk_umich_value.restype = ctypes.c_float
    lib.trk_umich_value.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double]
    return lib


def det_to_record(det, cam, width, height, pose_host, visible, opencv_return=0):
    """A detection dict of the synthetic sequence -> the fp32 pose record cp_decode_pnp would emit for it
    (first PnP: rep_mode 1 point assembly, through the host build of pose_core.h)."""
    r = np.zeros(L.CP_POSE_RECORD, np.float32)
    r[L.P_SCORE] = det["score"]
    r[L.P_CLS] = det["cls"]
    r[L.P_BBOX:L.P_BBOX + 4] = det["bbox"]
    r[L.P_CT:L.P_CT + 2] = det["ct"]
    for key, off, n in (("kps", L.P_KPS, 16), ("kps_displacement_mean", L.P_KPS_DISP_MEAN, 16),
                        ("kps_heatmap_mean", L.P_KPS_HM_MEAN, 16), ("kps_heatmap_std", L.P_KPS_HM_STD, 16),
                        ("kps_heatmap_height", L.P_KPS_HM_HEIGHT, 8), ("kps_displacement_std", L.P_KPS_DISP_STD, 16),
                        ("obj_scale", L.P_OBJ_SCALE, 3), ("obj_scale_uncertainty", L.P_OBJ_SCALE_UNC, 3),
                        ("tracking", L.P_TRACKING, 2), ("tracking_hp", L.P_TRACKING_HP, 16)):
        r[off:off + n] = np.asarray(det[key], np.float64)
    pts = np.ascontiguousarray(mg.assemble_points(det), np.float64)
    out = np.zeros(80, np.float64)
    st, npts = ctypes.c_int(0), ctypes.c_int(0)
    sc = np.ascontiguousarray(det["obj_scale"], np.float32)
    cam = np.ascontiguousarray(cam, np.float64)
    pose_host.host_solve_and_shell(pts.ctypes.data_as(ctypes.c_void_p), 16, sc.ctypes.data_as(ctypes.c_void_p),
                                   cam.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(width), ctypes.c_double(height),
                                   visible, opencv_return, out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(st),
                                   ctypes.byref(npts))
    r[L.P_STATUS] = st.value
    r[L.P_NPTS] = npts.value
    if st.value in (L.PNP_OK, L.PNP_INVISIBLE):
        r[L.P_LOCATION:L.P_LOCATION + 3] = out[0:3]
        r[L.P_QUAT:L.P_QUAT + 4] = out[3:7]
        r[L.P_REPROJ] = out[7]
        r[L.P_PROJ_CUBOID:L.P_PROJ_CUBOID + 16] = out[8:24]
        r[L.P_KPS_3D_CAM:L.P_KPS_3D_CAM + 27] = out[24:51]
        r[L.P_KPS_PNP:L.P_KPS_PNP + 18] = out[51:69]
    return r


def summarize_tracks(rows, n):
    """[T,320] track records -> the structure oracle.make_golden_tracker.summarize() produces."""
    out = {"n_boxes": int(sum(int(rows[i, L.T_IN_BOXES]) for i in range(n))), "tracks": []}
    for i in range(n):
        t = rows[i]
        e = {"tracking_id": int(t[L.T_ID]), "age": int(t[L.T_AGE]), "active": int(t[L.T_ACTIVE]), "score": float(t[L.P_SCORE]),
             "kps_mean_kf": t[L.T_KPS_MEAN_KF:L.T_KPS_MEAN_KF + 16].tolist(),
             "kps_std_kf": t[L.T_KPS_STD_KF:L.T_KPS_STD_KF + 16].tolist(),
             "obj_scale_kf": t[L.T_OBJ_SCALE_KF:L.T_OBJ_SCALE_KF + 3].tolist(),
             "obj_scale_uncertainty_kf": t[L.T_OBJ_SCALE_UNC_KF:L.T_OBJ_SCALE_UNC_KF + 3].tolist()}
        if int(t[L.P_STATUS]) in (L.PNP_OK, L.PNP_INVISIBLE):
            e["location"] = t[L.P_LOCATION:L.P_LOCATION + 3].tolist()
            e["quaternion_xyzw"] = t[L.P_QUAT:L.P_QUAT + 4].tolist()
        if int(t[L.T_PNP2_STATUS]) == L.PNP_OK:
            e["kps_pnp_kf"] = t[L.T_KPS_PNP_KF:L.T_KPS_PNP_KF + 18].tolist()
        out["tracks"].append(e)
    return out


def compare_to_golden(got, want, rel=2e-6, abs_px=2e-4):
    """fp32 records: values agree to ~1e-6 relative (2e-4 px on coordinates of a few hundred pixels)."""
    assert len(got) == len(want)
    for f, (g, w) in enumerate(zip(got, want)):
        assert g["n_boxes"] == w["n_boxes"], "frame %d: boxes %d vs %d" % (f, g["n_boxes"], w["n_boxes"])
        assert [t["tracking_id"] for t in g["tracks"]] == [t["tracking_id"] for t in w["tracks"]], "frame %d ids" % f
        for tg, tw in zip(g["tracks"], w["tracks"]):
            assert (tg["age"], tg["active"]) == (tw["age"], tw["active"]), (f, tw["tracking_id"])
            assert set(tg.keys()) == set(tw.keys()), (f, sorted(tg.keys()), sorted(tw.keys()))
            for k in tw:
                if isinstance(tw[k], list):
                    a, b = np.asarray(tg[k], np.float64), np.asarray(tw[k], np.float64)
                    if k == "quaternion_xyzw" and np.dot(a, b) < 0:
                        a = -a
                    atol = {"kps_mean_kf": abs_px, "quaternion_xyzw": 1e-5, "location": 1e-5}.get(k, 1e-6)
                    assert np.allclose(a, b, rtol=rel, atol=atol), "frame %d track %d %s: %g" % (
                        f, tw["tracking_id"], k, np.abs(a - b).max())


def test_track_core_matches_reference_golden(track_host, pose_host):
    gold = json.load(open(GOLD))
    o = gold["opt"]
    meta, frames = mg.make_sequence()
    cam = np.ascontiguousarray(meta["camera_matrix"], np.float64)
    h = track_host.trk_create(int(o["kalman"]), int(o["scale_pool"]), int(o["use_pnp"]), int(o["hps_uncertainty"]),
                              int(o["max_age"]), float(o["new_thresh"]), float(o["R"]), float(o["conf_border"][0]),
                              float(o["conf_border"][1]), VISIBLE[o["c"]], int(o["show_axes"]), 128)
    got = []
    out = np.zeros((128, L.CP_TRACK_RECORD), np.float32)
    for dets in frames:
        recs = np.stack([det_to_record(d, cam, meta["width"], meta["height"], pose_host, VISIBLE[o["c"]]) for d in dets])
        recs = np.ascontiguousarray(recs, np.float32)
        n = track_host.trk_step(h, recs.ctypes.data_as(ctypes.c_void_p), recs.shape[0], cam.ctypes.data_as(ctypes.c_void_p),
                                float(meta["width"]), float(meta["height"]), out.ctypes.data_as(ctypes.c_void_p))
        got.append(summarize_tracks(out.copy(), n))
    track_host.trk_destroy(h)
    compare_to_golden(got, gold["frames"])


def test_heatmap_primitives_match_reference_formulas(track_host):
    """gaussian_radius / draw_umich_gaussian values (utils/image.py:102-150), restated here in numpy."""
    def gaussian_radius(height, width, mo=0.7):
        b1 = height + width
        c1 = width * height * (1 - mo) / (1 + mo)
        r1 = (b1 + np.sqrt(b1 ** 2 - 4 * c1)) / 2
        b2 = 2 * (height + width)
        c2 = (1 - mo) * width * height
        r2 = (b2 + np.sqrt(b2 ** 2 - 16 * c2)) / 2
        a3 = 4 * mo
        b3 = -2 * mo * (height + width)
        c3 = (mo - 1) * width * height
        r3 = (b3 + np.sqrt(b3 ** 2 - 4 * a3 * c3)) / 2
        return min(r1, r2, r3)
    for (hh, ww) in ((10, 20), (133, 57), (1, 1), (400, 380)):
        assert abs(track_host.trk_gaussian_radius(hh, ww) - gaussian_radius(hh, ww)) <= 1e-12 * max(1, hh + ww)
    for r, k in ((0, 1.0), (3, 0.7), (17, 0.31)):
        d = 2 * r + 1
        sigma = d / 6
        y, x = np.ogrid[-r:r + 1, -r:r + 1]
        g = np.exp(-(x * x + y * y) / (2 * sigma * sigma))
        g[g < np.finfo(g.dtype).eps * g.max()] = 0
        want = (g * k).astype(np.float32)
        got = np.array([[track_host.trk_umich_value(dx, dy, r, k) for dx in range(-r, r + 1)] for dy in range(-r, r + 1)],
                       np.float32)
        assert np.abs(got - want).max() <= 1e-7
