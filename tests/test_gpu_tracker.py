"""CenterPoseTrack on the device (SURVEY.md rows a-T / f-2): cp_tracker_step vs the UNMODIFIED reference tracker
(tests/golden/tracker_seq.json), cp_tracker_render vs the unmodified `_get_additional_inputs`
(tests/golden/track_render.npz), and the tracking path of `ObjectPoseDetector.run()` / `run_batch(track=True)`."""
import json
import os
import types

import numpy as np
import pytest
import torch

import centerpose_b200 as cpb
from centerpose_b200 import _lib as L
from centerpose_b200 import synth
from oracle import make_golden_tracker as mg
from tests.test_track_core_host import GOLD, VISIBLE, compare_to_golden, det_to_record, summarize_tracks

pytestmark = pytest.mark.gpu
RENDER = os.path.join(os.path.dirname(GOLD), "track_render.npz")


def _opt_from_gold(o):
    opt = cpb.default_opt("dla_34", tracking_task=True, c=o["c"])
    opt.kalman, opt.scale_pool, opt.use_pnp = bool(o["kalman"]), bool(o["scale_pool"]), bool(o["use_pnp"])
    opt.hps_uncertainty, opt.max_age, opt.new_thresh, opt.R = bool(o["hps_uncertainty"]), int(o["max_age"]), o["new_thresh"], o["R"]
    opt.conf_border = {o["c"]: o["conf_border"]}
    opt.show_axes = bool(o["show_axes"])
    return opt


def _frame_records(dets, meta, pose_host, cat):
    cam = np.ascontiguousarray(meta["camera_matrix"], np.float64)
    recs = np.stack([det_to_record(d, cam, meta["width"], meta["height"], pose_host, VISIBLE[cat]) for d in dets])
    K = 100
    buf = np.zeros((1, K, L.CP_POSE_RECORD), np.float32)
    buf[0, :recs.shape[0]] = recs
    return torch.from_numpy(buf).cuda(), torch.tensor([recs.shape[0]], dtype=torch.int32).cuda()


def test_tracker_step_matches_reference_golden(cplib, pose_host):
    gold = json.load(open(GOLD))
    opt = _opt_from_gold(gold["opt"])
    meta, frames = mg.make_sequence()
    trk = cpb.Tracker(opt, streams=1)
    metat = cpb.make_meta(1, np.array([256., 256.], np.float32), 512.0, meta["width"], meta["height"], meta["camera_matrix"]).cuda()
    got = []
    for dets in frames:
        poses, nv = _frame_records(dets, meta, pose_host, opt.c)
        tr, n = trk.step_records(poses, nv, metat)
        got.append(summarize_tracks(tr[0].cpu().numpy(), int(n[0])))
    compare_to_golden(got, gold["frames"])
    d = trk.tracks                       # the reference-shaped dict view of the last frame
    assert [t["tracking_id"] for t in d] == [t["tracking_id"] for t in gold["frames"][-1]["tracks"]]
    assert {"kps_mean_kf", "kps_std_kf", "obj_scale_kf", "kps_fusion_mean", "tracking_id", "age", "active"} <= set(d[0])


def test_tracker_streams_are_independent_and_resettable(cplib, pose_host):
    """Two streams fed the same sequence with an offset give the same per-stream answers; reset() forgets the ids."""
    gold = json.load(open(GOLD))
    opt = _opt_from_gold(gold["opt"])
    meta, frames = mg.make_sequence()
    trk = cpb.Tracker(opt, streams=2)
    metat = cpb.make_meta(2, np.array([256., 256.], np.float32), 512.0, meta["width"], meta["height"], meta["camera_matrix"]).cuda()
    recs = [_frame_records(d, meta, pose_host, opt.c) for d in frames]
    outs = []
    for f in range(1, len(frames)):       # stream 0 sees frames 1.., stream 1 sees frames 0..
        poses = torch.cat([recs[f][0], recs[f - 1][0]])
        nv = torch.cat([recs[f][1], recs[f - 1][1]])
        tr, n = trk.step_records(poses, nv, metat)
        outs.append((tr.cpu().numpy(), n.cpu().numpy()))
    want = gold["frames"]
    got1 = [summarize_tracks(o[0][1], int(o[1][1])) for o in outs]
    compare_to_golden(got1, want[:len(got1)])
    trk.reset()
    tr, n = trk.step_records(recs[0][0].repeat(2, 1, 1), recs[0][1].repeat(2), metat)
    assert [int(v) for v in tr[0, :int(n[0]), L.T_ID].cpu()] == [1, 2, 3]


@pytest.mark.parametrize("case", mg.RENDER_CASES)
def test_render_matches_reference_additional_inputs(case, cplib, pose_host):
    name, nframes, ih, iw = case
    z = np.load(RENDER)
    gold = json.load(open(GOLD))
    opt = _opt_from_gold(gold["opt"])
    opt.pre_thresh, opt.render_hm_mode, opt.render_hmhp_mode = float(z["pre_thresh"]), int(z["render_hm_mode"]), int(z["render_hmhp_mode"])
    meta, frames = mg.make_sequence()
    trk = cpb.Tracker(opt, streams=1)
    metat = cpb.make_meta(1, np.array([256., 256.], np.float32), 512.0, meta["width"], meta["height"], meta["camera_matrix"]).cuda()
    for dets in frames[:nframes]:
        poses, nv = _frame_records(dets, meta, pose_host, opt.c)
        trk.step_records(poses, nv, metat)
    hm, hm_hp = trk.render(metat, z[name + "_trans_input"], ih, iw)
    torch.cuda.synchronize()
    for key, got in (("_hm", hm[0].cpu().numpy()), ("_hm_hp", hm_hp[0].cpu().numpy())):
        if name + key in z.files:
            want = z[name + key]
            assert got.shape == want.shape
            # fp32 records in, float64 -> float32 heat out: identical patches; values to 1e-6
            assert ((got != 0) == (want != 0)).all(), "%s%s: support differs" % (name, key)
            assert np.abs(got - want).max() <= 2e-6
        else:
            sub, ssum, nnz = z[name + key + "_sub4"], z[name + key + "_sum"], z[name + key + "_nnz"]
            assert np.abs(got[:, ::4, ::4] - sub).max() <= 2e-6
            assert ((got != 0).sum(axis=(1, 2)) == nnz).all()
            assert np.allclose(got.astype(np.float64).sum(axis=(1, 2)), ssum, rtol=1e-5, atol=1e-4)


def _tracking_detector(seed=31):
    opt = cpb.default_opt("dla_34", tracking_task=True)
    m = cpb.create_model(opt.arch, opt.heads, opt.head_conv, opt)
    m.load_state_dict(synth.seeded_state_dict(m, seed=seed, offset_std=0.3, head_gain=6.0))
    return cpb.ObjectPoseDetector(opt, model=m), opt


def test_run_tracking_path_contract(cplib):
    """run() with opt.tracking_task: previous image + rendered heat maps feed the network, the tracker steps, the
    return dict keeps the reference's keys and `results` carries tracking ids across frames."""
    det, opt = _tracking_detector()
    cam = synth.default_camera(512, 512)
    frames = synth.synthetic_frames(3, 512, 512, seed=77)
    ids = []
    for f in range(3):
        ret = det.run(frames[f], meta_inp={"camera_matrix": cam})
        assert {"results", "boxes", "output", "tot", "load", "pre", "net", "dec", "post", "merge", "pnp", "track"} == set(ret)
        assert set(opt.heads) <= set(ret["output"])
        for d in ret["results"]:
            assert {"tracking_id", "age", "active", "kps_mean_kf", "kps_fusion_mean", "obj_scale_kf"} <= set(d)
        ids.append([d["tracking_id"] for d in ret["results"]])
        assert det.pre_images is not None
        out = det.last_dict_out
        assert len(out["objects"]) == len(ret["results"])
        if out["objects"]:
            assert {"tracking_id", "kps_mean_kf", "kps_std_kf", "obj_scale_kf", "kps_fusion_mean", "tracking", "tracking_hp",
                    "kps_displacement_std", "obj_scale_uncertainty"} <= set(out["objects"][0])
            json.dumps(out)                       # serialisable like base_detector.py:741
    det.reset_tracking()
    assert det.pre_images is None and det.tracker.tracks == []


def test_run_batch_tracking_matches_run(cplib):
    """B video streams through run_batch(track=True) == each stream alone through run() (K partition held fixed: the
    tracker thresholds make the comparison discrete, see tests/util.py no_splitk)."""
    from tests.util import no_splitk
    with no_splitk():
        _batch_tracking_vs_run()


def _batch_tracking_vs_run():
    det, opt = _tracking_detector()
    cam = synth.default_camera(512, 512)
    vids = [synth.synthetic_frames(3, 512, 512, seed=100 + v) for v in range(2)]
    per_stream = []
    for v in range(2):
        det.reset_tracking()
        rows = []
        for f in range(3):
            ret = det.run(vids[v][f], meta_inp={"camera_matrix": cam})
            rows.append([(d["tracking_id"], d["score"], np.asarray(d["kps_mean_kf"]).reshape(-1)) for d in ret["results"]])
        per_stream.append(rows)
    det.reset_tracking()
    for f in range(3):
        batch = np.stack([vids[0][f], vids[1][f]])
        tracks, nt = det.run_batch(batch, cam, track=True)
        for v in range(2):
            want = per_stream[v][f]
            assert int(nt[v]) == len(want)
            for i, (tid, score, kf) in enumerate(want):
                assert int(tracks[v, i, L.T_ID]) == tid
                assert abs(float(tracks[v, i, L.P_SCORE]) - score) <= 1e-5
                assert np.abs(tracks[v, i, L.T_KPS_MEAN_KF:L.T_KPS_MEAN_KF + 16] - kf).max() <= 2e-2
