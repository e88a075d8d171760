"""`ObjectPoseDetector` on the B200-native hot path -- the drop-in for
/root/reference/src/lib/detectors/{base_detector,object_pose,detector_factory}.py.

`run()` keeps the reference's one-image semantics and its 12-key return dict
(base_detector.py:390-772): load -> pre_process (cv2, identical to the
reference) -> network -> decode -> post_process -> merge -> PnP.  Everything
after pre_process runs in libcenterpose_b200.so; the only host work left is
rebuilding the reference's Python result structures from the fixed-shape pose
records (`records_to_results`).  `run_batch()` is the batched entry point the
reference does not have (B = 32 / 256 configurations of BASELINE.json).
"""
import json
import os
import time

import numpy as np
import torch

from . import _lib
from .engine import decode_params, decode_pnp, make_meta, preprocess
from .model import create_model, load_model
from .tracker import Tracker, tracks_to_results


# ----------------------------------------------------------------------------- records -> reference structures
def record_to_result(rec):
    """One CP_POSE_RECORD row -> the reference's per-detection dict
    (post_process.py:27-63 + cuboid_pnp_shell.py:27-54)."""
    L = _lib
    r = np.asarray(rec, np.float32)
    d = {
        "score": float(r[L.P_SCORE]),
        "cls": int(r[L.P_CLS]),
        "obj_scale": r[L.P_OBJ_SCALE:L.P_OBJ_SCALE + 3].copy(),
        "obj_scale_uncertainty": r[L.P_OBJ_SCALE_UNC:L.P_OBJ_SCALE_UNC + 3].copy(),
        "kps_displacement_std": r[L.P_KPS_DISP_STD:L.P_KPS_DISP_STD + 16].copy(),
        "bbox": r[L.P_BBOX:L.P_BBOX + 4].astype(np.float64),
        "ct": [float(r[L.P_CT]), float(r[L.P_CT + 1])],
        "kps": r[L.P_KPS:L.P_KPS + 16].astype(np.float64),
        "tracking": r[L.P_TRACKING:L.P_TRACKING + 2].copy(),
        "tracking_hp": r[L.P_TRACKING_HP:L.P_TRACKING_HP + 16].copy(),
        "kps_displacement_mean": r[L.P_KPS_DISP_MEAN:L.P_KPS_DISP_MEAN + 16].astype(np.float64),
        "kps_heatmap_mean": r[L.P_KPS_HM_MEAN:L.P_KPS_HM_MEAN + 16].astype(np.float64),
        "kps_heatmap_std": r[L.P_KPS_HM_STD:L.P_KPS_HM_STD + 16].copy(),
        "kps_heatmap_height": r[L.P_KPS_HM_HEIGHT:L.P_KPS_HM_HEIGHT + 8].copy(),
    }
    st = int(r[L.P_STATUS])
    d["pnp_status"] = st
    if st in (L.PNP_OK, L.PNP_INVISIBLE):
        d["location"] = [float(v) for v in r[L.P_LOCATION:L.P_LOCATION + 3]]
        d["quaternion_xyzw"] = r[L.P_QUAT:L.P_QUAT + 4].astype(np.float64)
        d["projected_cuboid"] = r[L.P_PROJ_CUBOID:L.P_PROJ_CUBOID + 16].astype(np.float64).reshape(8, 2)
        d["kps_3d_cam"] = r[L.P_KPS_3D_CAM:L.P_KPS_3D_CAM + 27].astype(np.float64).reshape(9, 3)
        d["kps_pnp"] = r[L.P_KPS_PNP:L.P_KPS_PNP + 18].astype(np.float64).reshape(9, 2)
        d["reprojection_error"] = float(r[L.P_REPROJ])
    return d


def records_to_results(poses, n_valid, width, height):
    """poses [K,192], n_valid -> (results ndarray-of-dicts, boxes list) exactly
    shaped like base_detector.py:498,548-654."""
    res = [record_to_result(poses[i]) for i in range(int(n_valid))]
    boxes = []
    for d in res:
        if d["pnp_status"] == _lib.PNP_OK:
            kp = d["kps"].reshape(-1, 2)
            po = np.vstack([kp.mean(0, keepdims=True), kp]).copy()
            po[:, 0] /= width
            po[:, 1] /= height
            boxes.append((d["kps_pnp"], d["kps_3d_cam"], np.array(d["obj_scale"]), po, d))
    return np.array(res, dtype=object), boxes


def dets_to_dict(dets):
    """dets [B,K,128] -> the 13 arrays of decode.py:348-361."""
    L = _lib
    a = np.asarray(dets, np.float32)
    f = lambda off, n: a[..., off:off + n].copy()
    return {
        "bboxes": f(L.D_BBOX, 4), "scores": f(L.D_SCORE, 1), "kps": f(L.D_KPS, 16), "clses": f(L.D_CLS, 1),
        "obj_scale": f(L.D_OBJ_SCALE, 3), "obj_scale_uncertainty": f(L.D_OBJ_SCALE_UNC, 3),
        "tracking": f(L.D_TRACKING, 2), "tracking_hp": f(L.D_TRACKING_HP, 16),
        "kps_displacement_mean": f(L.D_KPS_DISP_MEAN, 16), "kps_displacement_std": f(L.D_KPS_DISP_STD, 16),
        "kps_heatmap_mean": f(L.D_KPS_HM_MEAN, 16), "kps_heatmap_std": f(L.D_KPS_HM_STD, 16),
        "kps_heatmap_height": f(L.D_KPS_HM_HEIGHT, 8),
    }


def affine_from_center_scale(c, s, out_w, out_h, inv=False):
    """The rot=0 case of utils/image.py:35-68: three float32 control points
    (centre, centre + (0, -s/2), and the perpendicular third point) handed to
    cv2.getAffineTransform, which is what the reference calls."""
    import cv2
    src = np.zeros((3, 2), np.float32)
    dst = np.zeros((3, 2), np.float32)
    src[0] = c
    src[1] = np.asarray(c, np.float32) + np.array([0, s * -0.5], np.float32)
    dst[0] = [out_w * 0.5, out_h * 0.5]
    dst[1] = np.array([out_w * 0.5, out_h * 0.5], np.float32) + np.array([0, out_w * -0.5], np.float32)
    for p in (src, dst):
        d = p[0] - p[1]
        p[2] = p[1] + np.array([-d[1], d[0]], np.float32)
    return cv2.getAffineTransform(dst, src) if inv else cv2.getAffineTransform(src, dst)


# ----------------------------------------------------------------------------- detector
class ObjectPoseDetector(object):
    def __init__(self, opt, model=None):
        if opt.gpus[0] < 0:
            raise RuntimeError("centerpose_b200 runs on a CUDA device only (--gpus -1 is the reference's CPU path)")
        opt.device = torch.device("cuda")
        self.opt = opt
        print("Creating model...")
        self.model = model if model is not None else create_model(opt.arch, opt.heads, opt.head_conv, opt)
        if getattr(opt, "load_model", ""):
            self.model = load_model(self.model, opt.load_model)
        self.model = self._to_device(self.model)
        self.model.eval()
        self.mean = np.array(opt.mean, dtype=np.float32).reshape(1, 1, 3)
        self.std = np.array(opt.std, dtype=np.float32).reshape(1, 1, 3)
        self.max_per_image = 100
        self.num_classes = opt.num_classes
        self.scales = opt.test_scales
        self.opt = opt
        self.pause = True
        self.pre_images = None
        self.tracker = None
        self.flip_idx = getattr(opt, "flip_idx", None)
        if getattr(opt, "refined_Kalman", False):
            raise NotImplementedError("opt.refined_Kalman (utils/tracker_baseline.py, CenterPose + Kalman baseline) is not "
                                      "on the accelerated path; use --tracking_task (CenterPoseTrack)")
        if getattr(opt, "tracking_task", False):
            # base_detector.py:53-54: the tracker state lives on the device (centerpose_b200/tracker.py)
            self.tracker = Tracker(opt, streams=1, device=opt.device)
        self._batch_tracker = None
        self._batch_pre = None

    def _to_device(self, t):
        """base_detector.py:41,436: everything the network touches lives on opt.device (always CUDA here)."""
        return t.to(self.opt.device)

    # base_detector.py:91-148 -- the reference's cv2 pre-processing (host side), all three modes
    def pre_process(self, image, scale, input_meta={}):
        import cv2
        height, width = image.shape[0:2]
        new_height = int(height * scale)
        new_width = int(width * scale)
        if self.opt.fix_short > 0:
            if height < width:
                inp_height = self.opt.fix_short
                inp_width = (int(width / height * self.opt.fix_short) + 63) // 64 * 64
            else:
                inp_height = (int(height / width * self.opt.fix_short) + 63) // 64 * 64
                inp_width = self.opt.fix_short
            c = np.array([width / 2, height / 2], dtype=np.float32)
            s = np.array([width, height], dtype=np.float32)
        elif self.opt.fix_res:
            inp_height, inp_width = self.opt.input_h, self.opt.input_w
            c = np.array([new_width / 2., new_height / 2.], dtype=np.float32)
            s = max(height, width) * 1.0
        else:
            inp_height = (new_height | self.opt.pad) + 1
            inp_width = (new_width | self.opt.pad) + 1
            c = np.array([new_width // 2, new_height // 2], dtype=np.float32)
            s = np.array([inp_width, inp_height], dtype=np.float32)
        s0 = float(s[0]) if isinstance(s, np.ndarray) else float(s)      # the affine only uses the width (image.py:44)
        trans_input = affine_from_center_scale(c, s0, inp_width, inp_height)
        out_height = inp_height // self.opt.down_ratio
        out_width = inp_width // self.opt.down_ratio
        trans_output = affine_from_center_scale(c, s0, out_width, out_height)
        resized = cv2.resize(image, (new_width, new_height))
        inp = cv2.warpAffine(resized, trans_input, (inp_width, inp_height), flags=cv2.INTER_LINEAR)
        inp = ((inp / 255. - self.mean) / self.std).astype(np.float32)
        images = torch.from_numpy(inp.transpose(2, 0, 1).reshape(1, 3, inp_height, inp_width))
        meta = {"c": c, "s": s, "height": height, "width": width, "out_height": out_height,
                "out_width": out_width, "inp_height": inp_height, "inp_width": inp_width,
                "trans_input": trans_input, "trans_output": trans_output}
        for k in ("pre_dets", "camera_matrix", "id"):
            if k in input_meta:
                meta[k] = input_meta[k]
        return images, meta

    def _meta_tensor(self, meta, batch=1):
        cam = meta.get("camera_matrix")
        if cam is None:
            if self.opt.use_pnp:
                raise ValueError("meta_inp['camera_matrix'] is required when opt.use_pnp is set (demo.py:141-147)")
            cam = np.eye(3)
        return make_meta(batch, meta["c"], meta["s"], meta["width"], meta["height"], cam)

    def process(self, images, pre_images=None, pre_hms=None, pre_hm_hp=None, pre_inds=None, return_time=False,
                meta=None, scale=1.0):
        """object_pose.py:131-165: network + sigmoid + decode.  Returns
        (output, dets[, forward_time]); the pose records of the fused stage are
        kept on `self._last` (host) / `self._last_dev` (device) for post_process / merge / PnP / tracking."""
        torch.cuda.synchronize()
        output = self.model(images, pre_images, pre_hms, pre_hm_hp)[-1]
        torch.cuda.synchronize()
        forward_time = time.time()
        prm = decode_params(self.opt, test_scale=float(scale))
        metat = self._meta_tensor(meta if meta is not None else self._dummy_meta(images), images.shape[0]).to(images.device)
        dets, poses, n_valid = decode_pnp(output, metat, prm, want_dets=True)
        output["hm"] = output["hm"].sigmoid_()
        if self.opt.hm_hp and not self.opt.mse_loss:
            output["hm_hp"] = output["hm_hp"].sigmoid_()
        output.update({"pre_inds": pre_inds})
        self._last_dev = (poses, n_valid, metat)
        self._last = (poses.cpu().numpy(), n_valid.cpu().numpy())
        dets = dets_to_dict(dets.cpu().numpy())
        if return_time:
            return output, dets, forward_time
        return output, dets

    def _dummy_meta(self, images):
        h, w = images.shape[2], images.shape[3]
        return {"c": np.array([w / 2., h / 2.], np.float32), "s": float(max(h, w)), "width": w, "height": h,
                "camera_matrix": np.eye(3)}

    def run(self, image_or_path_or_tensor, filename=None, meta_inp={}, preprocessed_flag=False):
        """base_detector.py:390-772.  One image per call, the reference's 12-key return dict.  After pre_process
        everything runs in libcenterpose_b200.so; post_process / merge_outputs / PnP are part of the fused decode call,
        so their stamps are 0 and `dec` carries the whole post-network stage."""
        import cv2
        load_time, pre_time, net_time, dec_time, post_time = 0, 0, 0, 0, 0
        merge_time, track_time, pnp_time, tot_time = 0, 0, 0, 0
        tracking = bool(getattr(self.opt, "tracking_task", False))
        if tracking and (len(self.scales) != 1 or self.scales[0] != 1.0):
            raise NotImplementedError("CenterPoseTrack runs at test_scales=[1] (the reference re-initialises the tracker "
                                      "state once per scale, base_detector.py:440-449)")
        start_time = time.time()
        pre_processed = preprocessed_flag
        if isinstance(image_or_path_or_tensor, np.ndarray):
            image = image_or_path_or_tensor
            if filename is not None:
                image_or_path_or_tensor = filename
        elif type(image_or_path_or_tensor) == type(""):
            image = cv2.imread(image_or_path_or_tensor)
        else:
            image = image_or_path_or_tensor["image"][0].numpy()
            pre_processed = True
        loaded_time = time.time()
        load_time += loaded_time - start_time

        # base_detector.py:421-497: one pass per test scale.  merge_outputs (object_pose.py:184-197) reads detections[0],
        # i.e. only the FIRST scale contributes results (with the soft-NMS forced on when several scales are listed); the
        # other passes still run because run() returns the `output` maps of the last one.
        first = None
        for si, scale in enumerate(self.scales):
            scale_start_time = time.time()
            if not pre_processed:
                images, meta = self.pre_process(image, scale, meta_inp)
            else:
                images = torch.from_numpy(np.expand_dims(image, axis=0))
                meta = meta_inp
            images = self._to_device(images)

            pre_hms, pre_hm_hp, pre_inds = None, None, None
            if tracking:
                if self.pre_images is None:                       # base_detector.py:444-449
                    print("Initialize tracking!")
                    self.pre_images = images
                    self.tracker.init_track(meta)
                if self.opt.pre_hm or self.opt.pre_hm_hp:         # :456-462, rendered on the device from the tracker state
                    if "trans_input" not in meta:
                        raise ValueError("tracking needs meta['trans_input'] (pre_process provides it)")
                    metat = self._meta_tensor(meta).to(images.device)
                    pre_hms, pre_hm_hp = self.tracker.render(metat, meta["trans_input"], images.shape[2], images.shape[3])
            torch.cuda.synchronize()
            pre_process_time = time.time()
            pre_time += pre_process_time - scale_start_time

            output, dets, forward_time = self.process(images, self.pre_images if tracking else None, pre_hms, pre_hm_hp,
                                                      pre_inds, return_time=True, meta=meta, scale=scale)
            torch.cuda.synchronize()
            net_time += forward_time - pre_process_time
            decode_time = time.time()
            dec_time += decode_time - forward_time
            if si == 0:
                first = (self._last, self._last_dev, meta)
        self._last, self._last_dev, meta = first

        # post_process + merge + PnP already happened inside cp_decode_pnp; unpack the records
        poses, n_valid = self._last
        results, boxes = records_to_results(poses[0], n_valid[0], meta["width"], meta["height"])
        if not self.opt.use_pnp:
            boxes = []
        post_process_time = time.time()
        post_time += post_process_time - decode_time
        pnp_process_time = post_process_time

        if tracking:                                          # :660-665 (gaussian_fusion :502-544 runs inside the step)
            pd, nd, md = self._last_dev
            self.tracker.step_records(pd, nd, md)
            rows, nt = self.tracker._host()
            results, boxes = tracks_to_results(rows[0], nt[0], meta["width"], meta["height"])
            if not self.opt.use_pnp:
                boxes = []
            self.pre_images = images
        end_time = time.time()
        track_time += end_time - pnp_process_time
        tot_time += end_time - start_time

        dict_out = self.build_dict_out(meta, results, boxes)
        self.last_dict_out = dict_out
        debug = int(getattr(self.opt, "debug", 0) or 0)
        if debug >= 1 and debug < 4:
            self.show_results(None, image, results)
        elif debug == 4:
            self.save_results(None, image, results, image_or_path_or_tensor, dict_out)
        elif debug == 6:
            self.save_results_eval(None, image, results, image_or_path_or_tensor, dict_out)
        return {"results": results, "boxes": boxes, "output": output, "tot": tot_time, "load": load_time,
                "pre": pre_time, "net": net_time, "dec": dec_time, "post": post_time, "merge": merge_time,
                "pnp": pnp_time, "track": track_time}

    # ------------------------------------------------------------------ result emitters
    def build_dict_out(self, meta, results, boxes):
        """The JSON payload of base_detector.py:672-754: camera matrix + one object per track (tracking) or per box."""
        opt = self.opt
        dict_out = {"camera_data": [], "objects": []}
        if "camera_matrix" in meta:
            dict_out["camera_data"] = np.asarray(meta["camera_matrix"]).tolist()
        lst = lambda v: np.asarray(v).tolist()      # noqa: E731
        if getattr(opt, "tracking_task", False):
            for tr in results:
                sc = np.asarray(tr["obj_scale"], np.float64)
                obj = {"class": opt.c, "ct": tr["ct"], "bbox": lst(tr["bbox"]), "confidence": tr["score"],
                       "kps_displacement_mean": lst(tr["kps_displacement_mean"]), "kps_heatmap_mean": lst(tr["kps_heatmap_mean"]),
                       "kps_heatmap_std": lst(tr["kps_heatmap_std"]), "kps_heatmap_height": lst(tr["kps_heatmap_height"]),
                       "obj_scale": (sc / sc[1]).tolist(), "tracking_id": tr["tracking_id"]}
                if opt.use_pnp:
                    if "location" in tr:
                        obj["location"] = tr["location"]
                        obj["quaternion_xyzw"] = lst(tr["quaternion_xyzw"])
                    if "kps_pnp" in tr:
                        obj["kps_pnp"] = lst(tr["kps_pnp"])
                        obj["kps_3d_cam"] = lst(tr["kps_3d_cam"])
                if getattr(opt, "obj_scale_uncertainty", False):
                    obj["obj_scale_uncertainty"] = lst(tr["obj_scale_uncertainty"])
                if getattr(opt, "kalman", False):
                    obj["kps_mean_kf"] = lst(tr["kps_mean_kf"])
                    obj["kps_std_kf"] = tr["kps_std_kf"]
                    if opt.use_pnp and "kps_pnp_kf" in tr:
                        obj["kps_pnp_kf"] = lst(tr["kps_pnp_kf"])
                        obj["kps_3d_cam_kf"] = lst(tr["kps_3d_cam_kf"])
                if getattr(opt, "scale_pool", False):
                    sk = np.asarray(tr["obj_scale_kf"], np.float64)
                    obj["obj_scale_kf"] = (sk / sk[1]).tolist()
                    obj["obj_scale_uncertainty_kf"] = lst(tr["obj_scale_uncertainty_kf"])
                if getattr(opt, "hps_uncertainty", False):
                    obj["kps_displacement_std"] = lst(tr["kps_displacement_std"])
                    obj["kps_fusion_mean"] = lst(tr["kps_fusion_mean"])
                    obj["kps_fusion_std"] = lst(tr["kps_fusion_std"])
                if getattr(opt, "tracking", False):
                    obj["tracking"] = lst(tr["tracking"])
                if getattr(opt, "tracking_hp", False):
                    obj["tracking_hp"] = lst(tr["tracking_hp"])
                dict_out["objects"].append(obj)
        else:
            for box in boxes:
                b = box[4]
                obj = {"class": opt.c, "ct": b["ct"], "bbox": lst(b["bbox"]), "confidence": b["score"],
                       "kps_displacement_mean": lst(b["kps_displacement_mean"]), "kps_heatmap_mean": lst(b["kps_heatmap_mean"]),
                       "kps_heatmap_std": lst(b["kps_heatmap_std"]), "kps_heatmap_height": lst(b["kps_heatmap_height"]),
                       "obj_scale": lst(b["obj_scale"])}
                if opt.use_pnp:
                    if "location" in b:
                        obj["location"] = b["location"]
                        obj["quaternion_xyzw"] = lst(b["quaternion_xyzw"])
                    if "kps_pnp" in b:
                        obj["kps_pnp"] = lst(b["kps_pnp"])
                        obj["kps_3d_cam"] = lst(b["kps_3d_cam"])
                dict_out["objects"].append(obj)
        return dict_out

    def _debugger(self, debugger):
        """The reference's Debugger (drawing) when its package is importable (drop-in use inside the reference tree);
        visualisation itself is outside the accelerated path, so without it only the JSON is written."""
        if debugger is not None:
            return debugger
        try:
            from lib.utils.debugger import Debugger
            return Debugger(dataset=self.opt.dataset, ipynb=(self.opt.debug == 3),
                            theme=getattr(self.opt, "debugger_theme", "white"))
        except Exception:
            return None

    def _draw(self, debugger, image, results, eval_mode=False):
        opt = self.opt
        debugger.add_img(image, img_id="out_img_pred")
        for bbox in results:
            if bbox["score"] > opt.vis_thresh and opt.reg_bbox:
                if getattr(opt, "tracking_task", False) and not eval_mode:
                    debugger.add_coco_bbox(bbox["bbox"], 0, bbox["score"], id=bbox["tracking_id"], img_id="out_img_pred")
                else:
                    debugger.add_coco_bbox(bbox["bbox"], 0, bbox["score"], img_id="out_img_pred")
                if "projected_cuboid" in bbox:
                    debugger.add_coco_hp(bbox["projected_cuboid"], img_id="out_img_pred", pred_flag="pnp")

    def show_results(self, debugger, image, results):
        """object_pose.py:280-317 (interactive display): needs the reference's Debugger."""
        dbg = self._debugger(debugger)
        if dbg is None:
            return
        self._draw(dbg, image, results)
        dbg.show_all_imgs(pause=self.pause)

    def save_results(self, debugger, image, results, image_or_path_or_tensor, dict_out=None):
        """object_pose.py:357-414: <demo_save>/<source name>/<frame>.json (+ the rendered image when a Debugger exists)."""
        opt = self.opt
        if os.path.isdir(opt.demo):
            target = os.path.join(opt.demo_save, os.path.basename(opt.demo))
        else:
            target = os.path.join(opt.demo_save, os.path.splitext(os.path.basename(opt.demo))[0])
        os.makedirs(target, exist_ok=True)
        dbg = self._debugger(debugger)
        if dbg is not None:
            self._draw(dbg, image, results)
            dbg.save_all_imgs_demo(image_or_path_or_tensor, path=target)
        if dict_out is not None:
            name = os.path.splitext(os.path.basename(image_or_path_or_tensor))[0]
            with open(os.path.join(target, name + ".json"), "w") as fp:
                json.dump(dict_out, fp)
            return os.path.join(target, name + ".json")

    def save_results_eval(self, debugger, image, results, image_or_path_or_tensor, dict_out=None, video_layout=False):
        """object_pose.py:319-355: demo/<checkpoint name>/[<video>/]<frame>.json, the layout the Objectron evaluator
        (tools/objectron_eval/eval_video_official.py:307-311) reads."""
        opt = self.opt
        if getattr(opt, "tracking_task", False) or getattr(opt, "eval_max_num", None) == 100:
            video_layout = True
        root = os.path.join("demo", os.path.splitext(os.path.basename(opt.load_model))[0])
        os.makedirs(root, exist_ok=True)
        key = image_or_path_or_tensor
        file_id = key[key.rfind("_") + 1:]
        folder = key[:key.rfind("_")]
        dbg = self._debugger(debugger)
        if dbg is not None:
            self._draw(dbg, image, results, eval_mode=True)
        if video_layout:
            vdir = os.path.join(root, folder)
            os.makedirs(vdir, exist_ok=True)
            if dbg is not None:
                dbg.save_all_imgs_eval(key, path=vdir, video_layout=True)
            path = os.path.join(vdir, file_id + ".json")
        else:
            if dbg is not None:
                dbg.save_all_imgs_eval(key, path=root, video_layout=False)
            path = os.path.join(root, "%s_%s.json" % (folder, file_id))
        if dict_out is not None:
            with open(path, "w") as fp:
                json.dump(dict_out, fp)
            return path

    # ------------------------------------------------------------------ batched API (not in the reference)
    def run_batch(self, frames, camera_matrix, pre_images=None, pre_hms=None, pre_hm_hp=None, to_host=True, track=False,
                  out=None):
        """frames: uint8 [B,H,W,3] (numpy / pinned CPU tensor / CUDA tensor) or a
        pre-processed fp32 [B,3,h,w] CUDA tensor.  One native cp_infer call for the
        whole batch.  Returns (poses [B,K,192], n_valid [B]) -- on the host when
        `to_host`, else as CUDA tensors.

        track=True (tracking models): the batch is B independent VIDEO STREAMS and every call is their next frame.
        The previous frames, the tracker state and the rendered previous-frame heat maps stay on the device; returns
        (tracks [B,T,320], n_tracks [B]) (layout: cp_track_field) instead.

        out: optional (poses, n_valid) CUDA tensors to write into (e.g. the views of a dist.PoseBuffer, so that the
        records land directly in the buffer of the all-gather / the pinned D2H copy)."""
        dev = self.opt.device
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(frames)
        if frames.dtype == torch.uint8:
            B, sh, sw, _ = frames.shape
            fr = frames.to(dev, non_blocking=True)
            x = preprocess(fr, self.opt.input_h, self.opt.input_w, self.opt.mean, self.opt.std)
            c, s = np.array([sw / 2., sh / 2.], np.float32), float(max(sh, sw))
            iw, ih = sw, sh
        else:
            x = frames.to(dev)
            B, _, ih, iw = x.shape
            c, s = np.array([iw / 2., ih / 2.], np.float32), float(max(ih, iw))
        # the per-frame meta (centre / scale / size / camera) of a fixed frame shape and camera is uploaded once
        cam_key = np.asarray(camera_matrix, np.float64).tobytes()
        mkey = (B, iw, ih, cam_key, str(dev))
        if getattr(self, "_meta_key", None) != mkey:
            self._meta_dev = make_meta(B, c, s, iw, ih, camera_matrix).to(dev)
            self._meta_key = mkey
        meta = self._meta_dev
        eng = self.model.engine(B, x.shape[2], x.shape[3], x.device)
        # the batched path pre-processes at scale 1 (what every shipped configuration uses); results of a multi-scale
        # opt are those of test_scales[0] (object_pose.py:188), which run() reproduces frame by frame
        if float(self.scales[0]) != 1.0:
            raise NotImplementedError("run_batch pre-processes at scale 1; use run() for test_scales[0] != 1")
        prm = decode_params(self.opt, test_scale=1.0)
        if track:
            if not getattr(self.opt, "tracking_task", False):
                raise ValueError("run_batch(track=True) needs a tracking model (opt.tracking_task)")
            trk = self._batch_tracker
            if trk is None or trk.streams != B:
                trk = self._batch_tracker = Tracker(self.opt, streams=B, device=x.device)
                self._batch_pre = None
            if self._batch_pre is None or self._batch_pre.shape != x.shape:
                self._batch_pre = x                                   # first frame: pre_images = images (base_detector.py:446)
            trans = affine_from_center_scale(c, s, x.shape[3], x.shape[2])
            pre_hms, pre_hm_hp = trk.render(meta, trans, x.shape[2], x.shape[3])
            _, poses, n_valid = eng.infer(x, meta, prm, self._batch_pre, pre_hms, pre_hm_hp)
            tracks, nt = trk.step_records(poses, n_valid, meta)
            self._batch_pre = x
            if to_host:
                return tracks.cpu().numpy(), nt.cpu().numpy()
            return tracks, nt
        _, poses, n_valid = eng.infer(x, meta, prm, pre_images, pre_hms, pre_hm_hp,
                                      poses=out[0] if out is not None else None, n_valid=out[1] if out is not None else None)
        if to_host:
            return poses.cpu().numpy(), n_valid.cpu().numpy()
        return poses, n_valid

    def reset_tracking(self):
        """base_detector.py:774-776."""
        if self.tracker is not None:
            self.tracker.reset()
        self.pre_images = None
        self._batch_pre = None
        if self._batch_tracker is not None:
            self._batch_tracker.reset()


detector_factory = {"object_pose": ObjectPoseDetector}
