"""Thin Python wrapper over the C ABI: one `Engine` = one native plan
(fixed arch / heads / resolution / max batch on one GPU) plus the output
buffers of the fused decode + PnP stage.  PyTorch only provides device
memory and the current stream here.
"""
import ctypes

import numpy as np
import torch

from . import _lib

HEAD_FIELDS = ("hm", "wh", "hps", "reg", "hm_hp", "hp_offset", "scale", "hps_uncertainty",
               "scale_uncertainty", "tracking", "tracking_hp")
_VISIBLE = {"book": 6, "chair": 6, "cereal_box": 6, "camera": 3, "bottle": 3, "cup": 3,
            "bike": 0, "laptop": 0, "shoe": 0}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def decode_params(opt=None, **over):
    """cp_decode_params from a reference-style `opt` (or keyword overrides)."""
    g = lambda n, d: over.get(n, getattr(opt, n, d) if opt is not None else d)
    p = _lib.CpDecodeParams()
    p.num_classes = int(g("num_classes", 1))
    p.num_joints = 8
    p.K = int(g("K", 100))
    p.rep_mode = int(g("rep_mode", 1))
    p.use_moments = int(bool(g("tracking_task", False)) or bool(g("refined_Kalman", False)))
    scales = list(g("test_scales", [1.0]))
    p.nms = int(bool(g("nms", True)))
    # object_pose.py:188-193: merge_outputs reads detections[0] (the first scale) and forces the soft-NMS when several
    # scales were requested; `test_scale` is the scale of the pass being decoded (detector.run passes it per pass)
    p.num_scales = len(scales)
    p.test_scale = float(over.get("test_scale", scales[0] if scales else 1.0))
    cat = g("c", "chair")
    if cat not in _VISIBLE:
        raise ValueError("unknown category '%s' (cuboid_pnp_shell.py:59-66)" % cat)
    p.visible_thresh = _VISIBLE[cat]
    p.opencv_return = int(bool(g("show_axes", False)))
    # object_pose.py:136-138: hm always goes through sigmoid_, hm_hp only when not opt.mse_loss
    p.apply_sigmoid = int(over.get("apply_sigmoid", 2 if bool(g("mse_loss", False)) else 1))
    p.use_pnp = int(bool(g("use_pnp", True)))
    p.vis_thresh = float(g("vis_thresh", 0.3))
    bal = g("balance_coefficient", None)
    p.balance = float(bal[cat]) if isinstance(bal, dict) else float(bal if bal is not None else 2.0)
    # 0 = the reference as pinned (torch==1.1.0 integer adds), 1 = the reference on torch >= 1.2
    p.modern_bool_semantics = int(bool(g("modern_bool_semantics", False)))
    if "use_moments" in over:
        p.use_moments = int(over["use_moments"])
    if "visible_thresh" in over:
        p.visible_thresh = int(over["visible_thresh"])
    return p


def make_meta(batch, c, s, img_w, img_h, cam, device=None, out=None):
    """[B,16] float64 meta rows (see centerpose_b200.h).  Scalars broadcast over the batch."""
    m = np.zeros((batch, _lib.CP_META_DOUBLES), np.float64)
    c = np.broadcast_to(np.asarray(c, np.float64).reshape(-1, 2), (batch, 2))
    m[:, 0:2] = c
    s_arr = np.asarray(s, np.float64)
    if s_arr.ndim == 0:
        m[:, 2] = float(s_arr)                 # one scalar for the whole batch
    elif s_arr.ndim == 1:
        m[:, 2] = np.broadcast_to(s_arr, (batch,))   # per-image scalar
    else:
        m[:, 2] = s_arr.reshape(batch, -1)[:, 0]     # [B,2] (w,h) pairs: the affine uses the width only
    m[:, 3] = img_w
    m[:, 4] = img_h
    m[:, 5:14] = np.broadcast_to(np.asarray(cam, np.float64).reshape(-1, 9), (batch, 9))
    t = torch.from_numpy(m)
    if device is not None:
        t = t.to(device)
    return t


def decode_pnp(heads, meta, prm, want_dets=True):
    """Run cp_decode_pnp on a dict of NCHW fp32 CUDA head tensors.
    Returns (dets [B,K,128] or None, poses [B,K,192], n_valid [B]) device tensors."""
    L = _lib.load()
    hm = heads["hm"]
    if not hm.is_cuda:
        raise RuntimeError("centerpose_b200.decode_pnp needs CUDA tensors (no CPU fallback)")
    B, _, H, W = hm.shape
    prm.batch, prm.out_h, prm.out_w = B, H, W
    hs = _lib.CpHeads()
    keep = []
    for f in HEAD_FIELDS:
        t = heads.get(f)
        if t is not None:
            t = t.contiguous().float()
            keep.append(t)
            setattr(hs, f, t.data_ptr())
    dev = hm.device
    meta = meta.to(dev, torch.float64).contiguous()
    dets = torch.empty((B, prm.K, _lib.CP_DETS_RECORD), dtype=torch.float32, device=dev) if want_dets else None
    poses = torch.empty((B, prm.K, _lib.CP_POSE_RECORD), dtype=torch.float32, device=dev)
    n_valid = torch.empty((B,), dtype=torch.int32, device=dev)
    nbytes = L.cp_decode_workspace_bytes(ctypes.byref(prm))
    if nbytes == 0:
        _lib.check(-1, "cp_decode_workspace_bytes")
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = L.cp_decode_pnp(ctypes.byref(prm), ctypes.byref(hs), _ptr(meta), _ptr(dets), _ptr(poses),
                             _ptr(n_valid), _ptr(ws), ctypes.c_size_t(nbytes), _stream())
    _lib.check(rc, "cp_decode_pnp")
    # the workspace must outlive the kernels; tie it to the outputs
    poses._cp_keep = (ws, keep, meta)
    return dets, poses, n_valid


class Engine(object):
    def __init__(self, arch, heads, head_conv, max_batch, height, width, device_index, tracking=False,
                 tracking_task_gru=False, precision="fp32"):
        self.L = _lib.load()
        self.heads = dict(heads)
        self.head_names = list(heads.keys())
        self.max_batch, self.height, self.width = int(max_batch), int(height), int(width)
        self.device = torch.device("cuda", device_index)
        self.tracking = bool(tracking)
        cfg = _lib.CpConfig()
        cfg.arch = {"dla_34": _lib.CP_ARCH_DLA34, "dlav1_34": _lib.CP_ARCH_DLAV1_34}[arch]
        cfg.tracking = int(tracking)
        cfg.tracking_task_gru = int(tracking_task_gru)
        cfg.max_batch, cfg.height, cfg.width = self.max_batch, self.height, self.width
        cfg.precision = _lib.PRECISIONS[precision]
        cfg.device = device_index
        cfg.head_conv = int(head_conv)
        cfg.num_heads = len(self.head_names)
        self._names = [n.encode() for n in self.head_names]
        for i, n in enumerate(self._names):
            cfg.head_names[i] = n
            cfg.head_channels[i] = int(self.heads[self.head_names[i]])
        self._cfg = cfg
        plan = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.L.cp_plan_create(ctypes.byref(cfg), ctypes.byref(plan)), "cp_plan_create")
        self.plan = plan
        self.weights_sig = None
        self.forward_launches = int(self.L.cp_plan_forward_launches(plan))
        self.plan_bytes = int(self.L.cp_plan_bytes(plan))

    def close(self):
        if getattr(self, "plan", None):
            self.L.cp_plan_destroy(self.plan)
            self.plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, sd):
        names, ptrs, numel, keep = [], [], [], []
        for k, v in sd.items():
            if not torch.is_tensor(v) or not v.dtype.is_floating_point:
                continue
            t = v.detach().to(self.device, torch.float32).contiguous()
            keep.append(t)
            names.append(k.encode())
            ptrs.append(t.data_ptr())
            numel.append(t.numel())
        n = len(names)
        a_names = (ctypes.c_char_p * n)(*names)
        a_ptrs = (ctypes.c_void_p * n)(*ptrs)
        a_numel = (ctypes.c_int64 * n)(*numel)
        with torch.cuda.device(self.device):
            rc = self.L.cp_plan_load_weights(self.plan, a_names, a_ptrs, a_numel, n, _stream())
            _lib.check(rc, "cp_plan_load_weights")
            torch.cuda.current_stream().synchronize()   # borrowed tensors may be freed after this

    def _dev(self, t, name, shape):
        """Every tensor handed to the kernels by data_ptr() must be a contiguous fp32 tensor ON THE PLAN'S DEVICE: a
        CPU tensor or one on another GPU would become a wild device pointer (sticky illegal-address error).  Like the
        reference (`images.to(opt.device)`, base_detector.py:436-441) tensors that live elsewhere are moved."""
        if not torch.is_tensor(t):
            raise TypeError("%s must be a torch tensor" % name)
        if tuple(t.shape) != tuple(shape):
            raise ValueError("%s %s does not fit the plan: expected %s" % (name, tuple(t.shape), tuple(shape)))
        if t.device != self.device or t.dtype != torch.float32:
            t = t.to(self.device, torch.float32)
        return t.contiguous()

    def _check_inputs(self, x, pre_img, pre_hm, pre_hm_hp):
        B = x.shape[0]
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, self.height, self.width) or B > self.max_batch or B < 1:
            raise ValueError("input %s does not fit the plan (%d,3,%d,%d)" % (tuple(x.shape), self.max_batch,
                                                                            self.height, self.width))
        xs = [self._dev(x, "images", (B, 3, self.height, self.width))]
        if self.tracking:
            if pre_img is None or pre_hm is None or pre_hm_hp is None:
                raise ValueError("a tracking plan needs pre_img, pre_hm and pre_hm_hp")
            xs += [self._dev(pre_img, "pre_img", (B, 3, self.height, self.width)),
                   self._dev(pre_hm, "pre_hm", (B, 1, self.height, self.width)),
                   self._dev(pre_hm_hp, "pre_hm_hp", (B, 8, self.height, self.width))]
        else:
            xs += [None, None, None]
        return B, xs

    def _out_tensor(self, t, name, shape, dtype):
        if t.device != self.device or t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
            raise ValueError("%s must be a contiguous %s tensor of shape %s on %s" % (name, dtype, tuple(shape), self.device))
        return t

    def forward(self, x, pre_img=None, pre_hm=None, pre_hm_hp=None, out=None):
        """Head logits {name: [B,C,H/4,W/4] fp32 CUDA}."""
        B, xs = self._check_inputs(x, pre_img, pre_hm, pre_hm_hp)
        if out is None:
            out = {n: torch.empty((B, c, self.height // 4, self.width // 4), dtype=torch.float32, device=self.device)
                   for n, c in self.heads.items()}
        else:
            for n, c in self.heads.items():
                self._out_tensor(out[n], "out[%s]" % n, (B, c, self.height // 4, self.width // 4), torch.float32)
        hp = (ctypes.c_void_p * len(self.head_names))(*[out[n].data_ptr() for n in self.head_names])
        with torch.cuda.device(self.device):
            rc = self.L.cp_forward(self.plan, B, _ptr(xs[0]), _ptr(xs[1]), _ptr(xs[2]), _ptr(xs[3]), hp, _stream())
        _lib.check(rc, "cp_forward")
        return out

    def profile(self, x, pre_img=None, pre_hm=None, pre_hm_hp=None):
        """One forward with CUDA events between the ops: list of dicts(name, kind, ms, flops, bytes)."""
        B, xs = self._check_inputs(x, pre_img, pre_hm, pre_hm_hp)
        out = {n: torch.empty((B, c, self.height // 4, self.width // 4), dtype=torch.float32, device=self.device)
               for n, c in self.heads.items()}
        hp = (ctypes.c_void_p * len(self.head_names))(*[out[n].data_ptr() for n in self.head_names])
        n_ops = int(self.L.cp_plan_num_ops(self.plan))
        stats = (_lib.CpOpStat * n_ops)()
        n = ctypes.c_int32(0)
        with torch.cuda.device(self.device):
            rc = self.L.cp_plan_profile(self.plan, B, _ptr(xs[0]), _ptr(xs[1]), _ptr(xs[2]), _ptr(xs[3]), hp, _stream(),
                                        stats, n_ops, ctypes.byref(n))
        _lib.check(rc, "cp_plan_profile")
        return [dict(name=stats[i].name.decode(), kind=int(stats[i].kind), ms=float(stats[i].ms),
                     flops=float(stats[i].flops), bytes=float(stats[i].bytes)) for i in range(n.value)]

    def infer(self, x, meta, prm, pre_img=None, pre_hm=None, pre_hm_hp=None, heads_out=None, want_dets=False,
              poses=None, n_valid=None, dets=None):
        """forward + decode + PnP in one native call.  Returns (dets|None, poses, n_valid)."""
        B, xs = self._check_inputs(x, pre_img, pre_hm, pre_hm_hp)
        if not torch.is_tensor(meta) or tuple(meta.shape) != (B, _lib.CP_META_DOUBLES):
            raise ValueError("meta must be a [%d,%d] tensor (make_meta)" % (B, _lib.CP_META_DOUBLES))
        if meta.device != self.device or meta.dtype != torch.float64 or not meta.is_contiguous():
            meta = meta.to(self.device, torch.float64).contiguous()
        if poses is None:
            poses = torch.empty((B, prm.K, _lib.CP_POSE_RECORD), dtype=torch.float32, device=self.device)
        else:
            self._out_tensor(poses, "poses", (B, prm.K, _lib.CP_POSE_RECORD), torch.float32)
        if n_valid is None:
            n_valid = torch.empty((B,), dtype=torch.int32, device=self.device)
        else:
            self._out_tensor(n_valid, "n_valid", (B,), torch.int32)
        if want_dets and dets is None:
            dets = torch.empty((B, prm.K, _lib.CP_DETS_RECORD), dtype=torch.float32, device=self.device)
        elif dets is not None:
            self._out_tensor(dets, "dets", (B, prm.K, _lib.CP_DETS_RECORD), torch.float32)
        hp = None
        if heads_out is not None:
            for n, c in self.heads.items():
                self._out_tensor(heads_out[n], "heads_out[%s]" % n, (B, c, self.height // 4, self.width // 4), torch.float32)
            hp = (ctypes.c_void_p * len(self.head_names))(*[heads_out[n].data_ptr() for n in self.head_names])
        with torch.cuda.device(self.device):
            rc = self.L.cp_infer(self.plan, B, _ptr(xs[0]), _ptr(xs[1]), _ptr(xs[2]), _ptr(xs[3]), ctypes.byref(prm),
                                 _ptr(meta), hp, _ptr(dets), _ptr(poses), _ptr(n_valid), _stream())
        _lib.check(rc, "cp_infer")
        return dets, poses, n_valid


class InferGraph(object):
    """cp_infer captured once in a CUDA graph (static buffers): a step is two small copies + one graph launch instead of
    ~80 kernel launches driven from the host -- what matters at batch 1, where the launch sequence, not the kernels,
    sets the latency.  `graph(x, meta)` -> (poses, n_valid) (views of the graph's own output buffers)."""

    def __init__(self, eng, batch, prm, tracking=False):
        self.eng, self.prm = eng, prm
        dev = eng.device
        H, W = eng.height, eng.width
        self.x = torch.zeros((batch, 3, H, W), dtype=torch.float32, device=dev)
        self.meta = torch.zeros((batch, _lib.CP_META_DOUBLES), dtype=torch.float64, device=dev)
        self.meta[:, 2] = float(max(H, W))
        self.meta[:, 3], self.meta[:, 4] = W, H
        self.meta[:, 5], self.meta[:, 9], self.meta[:, 13] = 1.0, 1.0, 1.0
        self.pre = None
        if tracking:
            self.pre = (torch.zeros_like(self.x), torch.zeros((batch, 1, H, W), dtype=torch.float32, device=dev),
                        torch.zeros((batch, 8, H, W), dtype=torch.float32, device=dev))
        self.poses = torch.zeros((batch, prm.K, _lib.CP_POSE_RECORD), dtype=torch.float32, device=dev)
        self.n_valid = torch.zeros((batch,), dtype=torch.int32, device=dev)
        pre = self.pre if self.pre is not None else (None, None, None)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):            # warm-up outside the capture: lazy allocations, launch attributes
            for _ in range(2):
                eng.infer(self.x, self.meta, prm, *pre, poses=self.poses, n_valid=self.n_valid)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            eng.infer(self.x, self.meta, prm, *pre, poses=self.poses, n_valid=self.n_valid)

    def __call__(self, x, meta, pre_img=None, pre_hm=None, pre_hm_hp=None):
        self.x.copy_(x, non_blocking=True)
        self.meta.copy_(meta, non_blocking=True)
        if self.pre is not None:
            for dst, src in zip(self.pre, (pre_img, pre_hm, pre_hm_hp)):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.poses, self.n_valid


def preprocess(frames_u8, dst_h, dst_w, mean, std, out=None, trans_input=None):
    """cp_preprocess: uint8 [B,H,W,3] CUDA -> fp32 [B,3,dst_h,dst_w] CUDA (bit-exact cv2.warpAffine + normalise).
    trans_input: optional 2x3 forward affine (meta['trans_input']); default = the fix_res affine of the frame size."""
    L = _lib.load()
    if not frames_u8.is_cuda or frames_u8.dtype != torch.uint8:
        raise RuntimeError("preprocess needs a uint8 CUDA tensor")
    B, sh, sw, _ = frames_u8.shape
    if out is None:
        out = torch.empty((B, 3, dst_h, dst_w), dtype=torch.float32, device=frames_u8.device)
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s = (ctypes.c_float * 3)(*[float(v) for v in std])
    with torch.cuda.device(frames_u8.device):
        if trans_input is not None:
            tm = (ctypes.c_double * 6)(*[float(v) for v in np.asarray(trans_input, np.float64).reshape(-1)])
            rc = L.cp_preprocess_affine(_ptr(frames_u8.contiguous()), _ptr(out), B, sh, sw, dst_h, dst_w, tm, m, s, _stream())
        else:
            rc = L.cp_preprocess(_ptr(frames_u8.contiguous()), _ptr(out), B, sh, sw, dst_h, dst_w, m, s, _stream())
    _lib.check(rc, "cp_preprocess")
    return out


def conv2d_nhwc(x, weight, bias=None, residual=None, stride=1, pad=0, relu=False, precision="fp32"):
    """cp_conv2d: x [B,H,W,Cin] NHWC fp32 CUDA, weight OIHW -> [B,Ho,Wo,Cout] NHWC."""
    L = _lib.load()
    if not x.is_cuda:
        raise RuntimeError("centerpose_b200 conv2d_nhwc needs CUDA tensors (no CPU fallback)")
    B, H, W, Cin = x.shape
    Cout, _, k, _ = weight.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    out = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32, device=x.device)
    ts = [t.contiguous().float() if t is not None else None for t in (x, weight, bias, residual)]
    with torch.cuda.device(x.device):
        rc = L.cp_conv2d(_ptr(ts[0]), _ptr(ts[1]), _ptr(ts[2]), _ptr(ts[3]), _ptr(out), B, H, W, Cin, Cout, k, stride,
                         pad, int(relu), _lib.PRECISIONS[precision], _stream())
    _lib.check(rc, "cp_conv2d")
    return out


def dcn_v2_forward(inp, weight, bias, offset, mask, kh=3, kw=3, sh=1, sw=1, ph=1, pw=1, dh=1, dw=1, dg=1,
                   precision="fp32"):
    """`_ext.dcn_v2_forward` signature (DCNv2/src/vision.cpp:4-9) on top of cp_dcn_v2_forward."""
    if (kh, kw, sh, sw, ph, pw, dh, dw, dg) != (3, 3, 1, 1, 1, 1, 1, 1, 1):
        raise RuntimeError("centerpose_b200 dcn_v2_forward: only 3x3 / stride 1 / pad 1 / dilation 1 / "
                           "deformable_group 1 is implemented (the configuration CenterPose uses)")
    if not inp.is_cuda:
        raise RuntimeError("centerpose_b200 dcn_v2_forward needs CUDA tensors (no CPU fallback)")
    L = _lib.load()
    B, C, H, W = inp.shape
    Co = weight.shape[0]
    out = torch.empty((B, Co, H, W), dtype=torch.float32, device=inp.device)
    ts = [t.contiguous().float() for t in (inp, weight, bias, offset, mask)]
    with torch.cuda.device(inp.device):
        rc = L.cp_dcn_v2_forward_ex(_ptr(ts[0]), _ptr(ts[1]), _ptr(ts[2]), _ptr(ts[3]), _ptr(ts[4]), _ptr(out),
                                    B, C, H, W, Co, _lib.PRECISIONS[precision], _stream())
    _lib.check(rc, "cp_dcn_v2_forward")
    out._cp_keep = ts
    return out


def dcn_v2_backward(inp, weight, bias, offset, mask, grad_output, kh=3, kw=3, sh=1, sw=1, ph=1, pw=1, dh=1, dw=1, dg=1,
                    precision="fp32"):
    """`_ext.dcn_v2_backward` signature (DCNv2/src/vision.cpp:11-17, dcn_v2.py:63-76) on top of cp_dcn_v2_backward.
    Returns (grad_input, grad_offset, grad_mask, grad_weight, grad_bias) as fresh tensors, like the reference."""
    if (kh, kw, sh, sw, ph, pw, dh, dw, dg) != (3, 3, 1, 1, 1, 1, 1, 1, 1):
        raise RuntimeError("centerpose_b200 dcn_v2_backward: only 3x3 / stride 1 / pad 1 / dilation 1 / "
                           "deformable_group 1 is implemented (the configuration CenterPose uses)")
    if not inp.is_cuda:
        raise RuntimeError("centerpose_b200 dcn_v2_backward needs CUDA tensors (no CPU fallback)")
    L = _lib.load()
    B, C, H, W = inp.shape
    Co = weight.shape[0]
    ts = [t.contiguous().float() for t in (inp, weight, offset, mask, grad_output)]
    for t, shape in zip(ts, ((B, C, H, W), (Co, C, 3, 3), (B, 18, H, W), (B, 9, H, W), (B, Co, H, W))):
        if tuple(t.shape) != shape or t.device != inp.device:
            raise ValueError("dcn_v2_backward: expected a %s tensor on %s, got %s on %s" % (shape, inp.device, tuple(t.shape), t.device))
    grads = [torch.empty_like(ts[0]), torch.empty_like(ts[2]), torch.empty_like(ts[3]), torch.empty_like(ts[1]),
             torch.empty((Co,), dtype=torch.float32, device=inp.device)]
    with torch.cuda.device(inp.device):
        rc = L.cp_dcn_v2_backward(*[_ptr(t) for t in ts], *[_ptr(g) for g in grads], B, C, H, W, Co,
                                  _lib.PRECISIONS[precision], _stream())
    _lib.check(rc, "cp_dcn_v2_backward")
    return tuple(grads)
