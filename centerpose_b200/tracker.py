"""`Tracker` -- the CenterPoseTrack state (/root/reference/src/lib/utils/tracker.py:15-302) on the device.

The reference keeps a Python list of dicts per video and runs association, a 32-state filterpy Kalman filter per
object, the scale pool and a second PnP on the host every frame; here the state of `streams` independent videos lives
in device memory and one native call (`cp_tracker_step`) advances all of them from the fixed-shape pose records that
`cp_decode_pnp` / `cp_infer` emit.  `cp_tracker_render` draws the previous-frame heat maps
(`BaseDetector._get_additional_inputs`, base_detector.py:150-388) straight into the network's `pre_hm` / `pre_hm_hp`
inputs.  The Python side only rebuilds the reference's dict structures for callers that want them (`tracks`).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .engine import _VISIBLE, _ptr, _stream


def _opt(opt, name, default):
    return getattr(opt, name, default)


def track_to_dict(row):
    """One CP_TRACK_RECORD row -> the reference's track dict (the keys run() / the debugger / the evaluator read)."""
    from .detector import record_to_result
    L = _lib
    t = np.asarray(row, np.float32)
    d = record_to_result(t[:L.CP_POSE_RECORD])
    d["tracking_id"] = int(t[L.T_ID])
    d["age"] = int(t[L.T_AGE])
    d["active"] = int(t[L.T_ACTIVE])
    d["kps_fusion_mean"] = t[L.T_KPS_FUSION_MEAN:L.T_KPS_FUSION_MEAN + 16].astype(np.float64)
    d["kps_fusion_std"] = t[L.T_KPS_FUSION_STD:L.T_KPS_FUSION_STD + 16].astype(np.float64)
    d["kps_mean_kf"] = t[L.T_KPS_MEAN_KF:L.T_KPS_MEAN_KF + 16].astype(np.float64).reshape(8, 2)
    d["kps_std_kf"] = [float(v) for v in t[L.T_KPS_STD_KF:L.T_KPS_STD_KF + 16]]
    d["obj_scale_kf"] = t[L.T_OBJ_SCALE_KF:L.T_OBJ_SCALE_KF + 3].astype(np.float64)
    d["obj_scale_uncertainty_kf"] = t[L.T_OBJ_SCALE_UNC_KF:L.T_OBJ_SCALE_UNC_KF + 3].astype(np.float64)
    d["pnp2_status"] = int(t[L.T_PNP2_STATUS])
    d["in_boxes"] = bool(int(t[L.T_IN_BOXES]))
    d["kps_conf_avg_kf"] = float(t[L.T_CONF_AVG])
    if d["pnp2_status"] == L.PNP_OK:
        d["kps_pnp_kf"] = t[L.T_KPS_PNP_KF:L.T_KPS_PNP_KF + 18].astype(np.float64).reshape(9, 2)
        d["kps_3d_cam_kf"] = t[L.T_KPS_3D_CAM_KF:L.T_KPS_3D_CAM_KF + 27].astype(np.float64).reshape(9, 3)
    return d


def tracks_to_results(rows, n, width, height):
    """[T,320] rows of one stream -> (results list of dicts, boxes list of tuples) shaped like Tracker.step's return
    (tracker.py:271-295): a box = (kps_pnp_kf, kps_3d_cam_kf, obj_scale, kps_ori_kf, track)."""
    res = [track_to_dict(rows[i]) for i in range(int(n))]
    boxes = []
    for d in res:
        if d["in_boxes"] and "kps_pnp_kf" in d:
            kp = np.asarray(d["kps"], np.float64).reshape(-1, 2)
            po = np.vstack([kp.mean(0, keepdims=True), kp]).copy()
            po[:, 0] /= width
            po[:, 1] /= height
            d["kps_ori_kf"] = po
            boxes.append((d["kps_pnp_kf"], d["kps_3d_cam_kf"], np.array(d["obj_scale"]), po, d))
    return res, boxes


class Tracker(object):
    def __init__(self, opt, streams=1, device=None, max_tracks=_lib.CP_MAX_K):
        if _opt(opt, "hungarian", False):
            raise NotImplementedError("centerpose_b200 Tracker implements the greedy association (opt.hungarian is off in demo.py)")
        self.L = _lib.load()
        self.opt = opt
        self.streams = int(streams)
        self.max_tracks = int(max_tracks)
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        cat = _opt(opt, "c", "chair")
        border = _opt(opt, "conf_border", {cat: [3, 9]})
        border = border[cat] if isinstance(border, dict) else border
        cfg = _lib.CpTrackerConfig()
        cfg.streams, cfg.max_tracks = self.streams, self.max_tracks
        cfg.kalman = int(bool(_opt(opt, "kalman", True)))
        cfg.scale_pool = int(bool(_opt(opt, "scale_pool", True)))
        cfg.use_pnp = int(bool(_opt(opt, "use_pnp", True)))
        cfg.hps_uncertainty = int(bool(_opt(opt, "hps_uncertainty", True)))
        cfg.max_age = int(_opt(opt, "max_age", 5))
        cfg.visible_thresh = _VISIBLE[cat]
        cfg.opencv_return = int(bool(_opt(opt, "show_axes", False)))
        cfg.render_hm_mode = int(_opt(opt, "render_hm_mode", 1))
        cfg.render_hmhp_mode = int(_opt(opt, "render_hmhp_mode", 2))
        cfg.device = self.device.index
        cfg.new_thresh = float(_opt(opt, "new_thresh", 0.3))
        cfg.pre_thresh = float(_opt(opt, "pre_thresh", -1))
        cfg.R = float(_opt(opt, "R", 20))
        cfg.conf_lo, cfg.conf_hi = float(border[0]), float(border[1])
        self._cfg = cfg
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.L.cp_tracker_create(ctypes.byref(cfg), ctypes.byref(h)), "cp_tracker_create")
        self._h = h
        self.meta = None
        self._rows = None          # host copy of the latest step: (rows [B,T,320], n [B])
        self._dev = None           # device tensors of the latest step
        self._dicts = None

    def close(self):
        if getattr(self, "_h", None):
            self.L.cp_tracker_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference surface -------------------------------------------------------------------------------------
    def reset(self, index=-1):
        """Tracker.reset (tracker.py:50-52)."""
        with torch.cuda.device(self.device):
            _lib.check(self.L.cp_tracker_reset(self._h, int(index), _stream()), "cp_tracker_reset")
        self._rows = self._dev = self._dicts = None

    def init_track(self, meta):
        """Tracker.init_track (tracker.py:22-48).  Seeding from meta['pre_dets'] (ground-truth experiments) is not
        on the accelerated path."""
        if meta is not None and "pre_dets" in meta and len(meta["pre_dets"]):
            raise NotImplementedError("seeding the device tracker from meta['pre_dets'] is not supported")
        self.meta = meta

    @property
    def tracks(self):
        """The reference's `self.tracker.tracks` of stream 0 (list of dicts), rebuilt from the latest step."""
        if self._dicts is None:
            if self._dev is None:
                return []
            rows, n = self._host()
            self._dicts = [track_to_dict(rows[0, i]) for i in range(int(n[0]))]
        return self._dicts

    # ---- device entry points -----------------------------------------------------------------------------------
    def _host(self):
        if self._rows is None:
            tr, n = self._dev
            self._rows = (tr.cpu().numpy(), n.cpu().numpy())
        return self._rows

    def step_records(self, poses, n_valid, meta, out=None):
        """poses [B,K,192] / n_valid [B] / meta [B,16] CUDA tensors (as cp_infer emits them) -> (tracks [B,T,320],
        n_tracks [B]) CUDA tensors.  Stream b of the tracker consumes poses[b]."""
        B, K, R = poses.shape
        if R != _lib.CP_POSE_RECORD or B > self.streams:
            raise ValueError("poses %s does not fit a tracker of %d streams" % (tuple(poses.shape), self.streams))
        for t, dt in ((poses, torch.float32), (n_valid, torch.int32), (meta, torch.float64)):
            if t.device != self.device or t.dtype != dt or not t.is_contiguous():
                raise ValueError("tracker inputs must be contiguous %s tensors on %s" % (dt, self.device))
        if out is None:
            out = (torch.empty((B, self.max_tracks, _lib.CP_TRACK_RECORD), dtype=torch.float32, device=self.device),
                   torch.empty((B,), dtype=torch.int32, device=self.device))
        with torch.cuda.device(self.device):
            rc = self.L.cp_tracker_step(self._h, B, _ptr(poses), _ptr(n_valid), K, _ptr(meta), _ptr(out[0]), _ptr(out[1]),
                                        _stream())
        _lib.check(rc, "cp_tracker_step")
        self._dev, self._rows, self._dicts = out, None, None
        return out

    def render(self, meta, trans_input, inp_h, inp_w, out=None):
        """Previous-frame heat maps of every stream: (pre_hm [B,1,h,w], pre_hm_hp [B,8,h,w]) fp32 CUDA."""
        B = meta.shape[0]
        tr = torch.as_tensor(np.asarray(trans_input, np.float64).reshape(-1, 6)) if not torch.is_tensor(trans_input) else trans_input
        if tr.shape[0] == 1 and B > 1:
            tr = tr.expand(B, 6)
        tr = tr.to(self.device, torch.float64).contiguous()
        meta = meta.to(self.device, torch.float64).contiguous()
        if out is None:
            out = (torch.empty((B, 1, inp_h, inp_w), dtype=torch.float32, device=self.device),
                   torch.empty((B, 8, inp_h, inp_w), dtype=torch.float32, device=self.device))
        with torch.cuda.device(self.device):
            rc = self.L.cp_tracker_render(self._h, B, _ptr(meta), _ptr(tr), int(inp_h), int(inp_w), _ptr(out[0]), _ptr(out[1]),
                                          _stream())
        _lib.check(rc, "cp_tracker_render")
        out[0]._cp_keep = (meta, tr)
        return out
