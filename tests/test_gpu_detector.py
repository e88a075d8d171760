"""End-to-end `ObjectPoseDetector.run()` / `run_batch()` on the GPU and the batched pre-process kernel."""
import numpy as np
import pytest
import torch

import centerpose_b200 as cpb
from centerpose_b200 import _lib as L
from centerpose_b200 import synth
from oracle import decode_ref
from tests.util import compare_records, oracle_records

pytestmark = pytest.mark.gpu

KEYS = {"results", "boxes", "output", "tot", "load", "pre", "net", "dec", "post", "merge", "pnp", "track"}


def _detector(head_gain=6.0, seed=21):
    opt = cpb.default_opt("dla_34")
    m = cpb.create_model(opt.arch, opt.heads, opt.head_conv, opt)
    m.load_state_dict(synth.seeded_state_dict(m, seed=seed, offset_std=1.0, head_gain=head_gain))
    return cpb.ObjectPoseDetector(opt, model=m), opt


def test_run_contract_and_consistency(cplib):
    det, opt = _detector()
    img = synth.synthetic_frames(1, 600, 800, seed=3)[0]
    cam = np.array([[663.0287679036459, 0, 300.2775065104167], [0, 663.0287679036459, 395.00066121419275], [0, 0, 1]])
    ret = det.run(img, meta_inp={"camera_matrix": cam})
    assert set(ret) == KEYS
    out = ret["output"]
    assert set(opt.heads) <= set(out) and out["hm"].shape == (1, 1, 128, 128)
    assert float(out["hm"].min()) >= 0 and float(out["hm"].max()) <= 1        # sigmoid applied like object_pose.py:136
    # the records run() unpacked must equal the oracle pipeline run on the heads the GPU produced
    heads = {k: out[k][0].cpu().numpy() for k in opt.heads}
    for k in ("hm", "hm_hp"):                                                   # back to logits for the oracle entry point
        p = np.clip(heads[k].astype(np.float64), 1e-12, 1 - 1e-7)
        heads[k] = np.log(p / (1 - p)).astype(np.float32)
    prm = decode_ref.DecodeParams(rep_mode=1, vis_thresh=opt.vis_thresh, category=opt.c)
    c = np.array([800 / 2., 600 / 2.], np.float32)
    _, want = oracle_records(heads, prm, cam, 800, 600, c, 800.0, L)
    assert len(ret["results"]) == want.shape[0]
    for d in ret["results"]:
        assert {"score", "cls", "obj_scale", "bbox", "ct", "kps", "kps_displacement_mean", "kps_heatmap_mean",
                "kps_heatmap_std", "kps_heatmap_height", "tracking", "tracking_hp"} <= set(d)
    for b in ret["boxes"]:
        assert b[0].shape == (9, 2) and b[1].shape == (9, 3) and b[3].shape == (9, 2) and "location" in b[4]


def test_run_multi_scale(cplib):
    """opt.test_scales = [0.75, 1.0] (base_detector.py:421-497, object_pose.py:167-197): run() returns the detections of
    the FIRST scale -- coordinates divided by 0.75, soft-NMS forced on -- and the head maps of the LAST pass."""
    det, opt = _detector(head_gain=1.0)
    img = synth.synthetic_frames(1, 600, 800, seed=3)[0]
    # (_detector's default weights saturate every score at 1.0: an all-ties scene.  Calibrate the heat-map biases so that
    # a handful of distinct peaks pass the thresholds -- setup only, as in bench.py)
    x0, _ = det.pre_process(img, 0.75, {})
    with torch.no_grad():
        synth.calibrate_head_bias(det.model, det.model(x0.cuda())[-1], target=6)
    opt.test_scales = [0.75, 1.0]
    opt.nms = False
    det.scales = opt.test_scales
    cam = np.array([[663.0287679036459, 0, 300.2775065104167], [0, 663.0287679036459, 395.00066121419275], [0, 0, 1]])
    ret = det.run(img, meta_inp={"camera_matrix": cam})
    assert set(ret) == KEYS
    # the heads of the 0.75 pass, recomputed through the public pieces, decoded by the CPU oracle at that scale
    images, meta = det.pre_process(img, 0.75, {"camera_matrix": cam})
    assert tuple(meta["c"]) == (300.0, 225.0) and meta["s"] == 800.0
    out, _ = det.process(images.cuda(), meta=meta, scale=0.75)
    heads = {k: out[k][0].cpu().numpy() for k in opt.heads}
    for k in ("hm", "hm_hp"):
        p = np.clip(heads[k].astype(np.float64), 1e-12, 1 - 1e-7)
        heads[k] = np.log(p / (1 - p)).astype(np.float32)
    prm = decode_ref.DecodeParams(rep_mode=1, vis_thresh=opt.vis_thresh, category=opt.c, nms=False, num_scales=2)
    _, want = oracle_records(heads, prm, cam, 800, 600, meta["c"], meta["s"], L, scale=0.75)
    assert len(ret["results"]) == want.shape[0] and want.shape[0] > 0
    # (the maps went through sigmoid and back, so soft-NMS scores agree to ~1e-4 and near-ties may swap places: match
    # every result to its oracle record instead of relying on the order)
    left = list(range(want.shape[0]))
    for d in ret["results"]:
        j = min(left, key=lambda q: np.abs(np.asarray(d["bbox"]) - want[q, L.P_BBOX:L.P_BBOX + 4]).max())
        left.remove(j)
        w = want[j]
        assert abs(d["score"] - w[L.P_SCORE]) <= 1e-3, (d["score"], w[L.P_SCORE])
        assert np.abs(np.asarray(d["bbox"]) - w[L.P_BBOX:L.P_BBOX + 4]).max() <= 0.05
        assert np.abs(np.asarray(d["kps"]) - w[L.P_KPS:L.P_KPS + 16]).max() <= 0.05
        assert np.abs(np.asarray(d["kps_displacement_mean"]) - w[L.P_KPS_DISP_MEAN:L.P_KPS_DISP_MEAN + 16]).max() <= 0.05
    # last pass = scale 1.0: the returned maps are those of a plain single-scale run
    opt.test_scales = [1.0]
    det.scales = opt.test_scales
    ret1 = det.run(img, meta_inp={"camera_matrix": cam})
    assert torch.equal(ret["output"]["hm"], ret1["output"]["hm"])
    with pytest.raises(NotImplementedError):
        opt.test_scales = [0.5]
        det.scales = opt.test_scales
        det.run_batch(synth.synthetic_frames(1, 512, 512, seed=1), cam)
    opt.test_scales = [1.0]
    det.scales = opt.test_scales


def test_run_batch_matches_run(cplib):
    """run() is run_batch() of one frame + unpacking: the same frame through both gives the same records, alone or inside
    a batch of four (K partition held fixed, see tests/util.py no_splitk)."""
    from tests.util import no_splitk
    det, opt = _detector()
    frames = synth.synthetic_frames(4, 512, 512, seed=9)
    cam = synth.default_camera(512, 512)
    with no_splitk():
        poses, n_valid = det.run_batch(frames, cam)
        assert poses.shape == (4, opt.K, L.CP_POSE_RECORD)
        for b in (0, 3):
            ret = det.run(frames[b], meta_inp={"camera_matrix": cam})
            assert len(ret["results"]) == n_valid[b]
            for i, d in enumerate(ret["results"]):       # same input bits, same K partition: the same records, bit for bit
                assert d["score"] == poses[b, i, L.P_SCORE]
                assert np.array_equal(np.asarray(d["kps"], np.float32), poses[b, i, L.P_KPS:L.P_KPS + 16])
                assert np.array_equal(np.asarray(d["bbox"], np.float32), poses[b, i, L.P_BBOX:L.P_BBOX + 4])
    # default plan (split-K on): one frame through run() and through run_batch()
    for b in (0, 3):
        ret = det.run(frames[b], meta_inp={"camera_matrix": cam})
        p1, n1 = det.run_batch(frames[b:b + 1], cam)
        assert len(ret["results"]) == n1[0]
        for i, d in enumerate(ret["results"]):
            assert abs(d["score"] - p1[0, i, L.P_SCORE]) <= 1e-4
            assert np.abs(d["kps"] - p1[0, i, L.P_KPS:L.P_KPS + 16]).max() <= 0.05


def test_preprocess_is_bit_exact(cplib):
    """Row f-1: the batched pre-process kernel reproduces cv2.warpAffine(INTER_LINEAR) + the normalisation of
    base_detector.py:128-134 BIT FOR BIT on the Objectron frame shapes (byte work: the bar is exact)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import preprocess_ref as pr
    det, opt = _detector()
    for (h, w) in ((512, 512), (600, 800), (480, 640), (800, 600), (375, 500)):
        fr = synth.synthetic_frames(2, h, w, seed=h)
        got = cpb.preprocess(torch.from_numpy(fr).cuda(), 512, 512, opt.mean, opt.std).cpu().numpy()
        for b in range(2):
            want, meta = det.pre_process(fr[b], 1.0)              # the reference's cv2 path
            assert np.array_equal(got[b], want[0].numpy()), (h, w, np.abs(got[b] - want[0].numpy()).max())
            assert np.array_equal(got[b], pr.pre_process(fr[b], 512, 512, opt.mean, opt.std)[0])
        # explicit trans_input (what run() computes) and an arbitrary rotated crop
        got2 = cpb.preprocess(torch.from_numpy(fr).cuda(), 512, 512, opt.mean, opt.std, trans_input=meta["trans_input"]).cpu().numpy()
        assert np.array_equal(got2[1], want[0].numpy())
    fr = synth.synthetic_frames(1, 300, 420, seed=2)
    M = cv2.getRotationMatrix2D((210, 150), 17.0, 0.8)
    got = cpb.preprocess(torch.from_numpy(fr).cuda(), 192, 256, opt.mean, opt.std, trans_input=M).cpu().numpy()
    inp = cv2.warpAffine(fr[0], M, (256, 192), flags=cv2.INTER_LINEAR)
    want = ((inp / 255. - det.mean) / det.std).astype(np.float32).transpose(2, 0, 1)
    assert np.array_equal(got[0], want)


def test_batch_pipeline_matches_run_batch(cplib):
    """centerpose_b200.BatchPipeline (double-buffered upload / compute / download) returns, batch for batch and in
    submission order, exactly what run_batch() returns -- with pinned and with pageable frames."""
    det, opt = _detector()
    cam = synth.default_camera(512, 512)
    batches = [synth.synthetic_frames(2, 512, 512, seed=40 + i) for i in range(5)]
    want = [det.run_batch(b, cam) for b in batches]
    pipe = cpb.BatchPipeline(det, 2, 512, 512, cam, depth=2)
    got = []
    for i, b in enumerate(batches):
        if pipe.in_flight == pipe.depth:
            p, n = pipe.collect()
            got.append((p.copy(), n.copy()))
        t = torch.from_numpy(b)
        pipe.submit(t.pin_memory() if i % 2 else t)
    with pytest.raises(RuntimeError):
        pipe.submit(batches[0]); pipe.submit(batches[0])
    while pipe.in_flight:
        p, n = pipe.collect()
        got.append((p.copy(), n.copy()))
    got = got[:5]
    assert len(got) == 5
    for (p, n), (wp, wn) in zip(got, want):
        assert (n == wn).all()
        assert np.array_equal(p, wp)
    with pytest.raises(ValueError):
        pipe.submit(np.zeros((2, 100, 100, 3), np.uint8))
