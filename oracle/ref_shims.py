"""TEST INFRASTRUCTURE ONLY -- import shims that let the UNMODIFIED reference
(`/root/reference/src`) run on a CPU-only box with the libraries in this image.

Nothing under `oracle/` is ever imported by the product path
(`centerpose_b200/`); only `tests/`, `__graft_entry__.smoke()`,
`oracle/make_golden.py` and `bench.py`'s cpu_baseline / `--impl reference`
legs may use it.

`/root/reference` exists only in the build container, never on the GPU box:
`reference_available()` gates every user of this file.

What is shimmed -- behaviour, not code:
  * `_ext.dcn_v2_forward(...)`  -> `torchvision.ops.deform_conv2d` (same
    offset / mask channel layout as DCNv2/src/cpu/dcn_v2_im2col_cpu.cpp:160-172)
  * `torch.utils.model_zoo.load_url` -> dummy dict (pose_dla_dcn.py:330 has no
    network here); keeps the `num_classes` probe at :331 working
  * `pyrr.Quaternion.from_axis_rotation`, `progress.bar.Bar`, `simplejson`,
    `matplotlib(.pyplot)`, `filterpy.*`, `sklearn.utils.linear_assignment_`
  * `torch.cuda.synchronize` -> no-op when CUDA is absent
    (base_detector.py:466 calls it unconditionally)
"""
import os
import sys
import types
import json

import numpy as np
import torch

REF_ROOT = os.environ.get("CENTERPOSE_REFERENCE", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")
REF_LIB = os.path.join(REF_SRC, "lib")

_installed = False


def reference_available():
    return os.path.isdir(REF_LIB)


def _dcn_v2_forward(inp, weight, bias, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, dg):
    import torchvision.ops as tvo
    return tvo.deform_conv2d(inp, offset, weight, bias, stride=(sh, sw),
                             padding=(ph, pw), dilation=(dh, dw), mask=mask)


def _not_impl(*a, **k):
    raise NotImplementedError("shimmed reference: backward / PSROI are not on the inference path")


def install():
    """Idempotently install the stub modules and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)

    ext = types.ModuleType("_ext")
    ext.dcn_v2_forward = _dcn_v2_forward
    ext.dcn_v2_backward = _not_impl
    ext.dcn_v2_psroi_pooling_forward = _not_impl
    ext.dcn_v2_psroi_pooling_backward = _not_impl
    sys.modules["_ext"] = ext

    # pyrr: only Quaternion.from_axis_rotation is used (cuboid_pnp_solver.py:247)
    pyrr = types.ModuleType("pyrr")

    class Quaternion(object):
        @staticmethod
        def from_axis_rotation(axis, theta):
            axis = np.asarray(axis, dtype=np.float64).reshape(3)
            theta = float(np.asarray(theta).reshape(-1)[0])
            half = theta * 0.5
            return np.array([axis[0] * np.sin(half), axis[1] * np.sin(half),
                             axis[2] * np.sin(half), np.cos(half)])
    pyrr.Quaternion = Quaternion
    sys.modules["pyrr"] = pyrr

    progress = types.ModuleType("progress")
    bar = types.ModuleType("progress.bar")

    class Bar(object):
        def __init__(self, *a, **k):
            pass

        def next(self):
            pass

        def finish(self):
            pass
    bar.Bar = Bar
    progress.bar = bar
    sys.modules["progress"] = progress
    sys.modules["progress.bar"] = bar

    sys.modules.setdefault("simplejson", json)

    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    mpl.use = lambda *a, **k: None
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)

    fp = types.ModuleType("filterpy")
    fpk = types.ModuleType("filterpy.kalman")
    fpc = types.ModuleType("filterpy.common")

    class KalmanFilter(object):
        """Functional stand-in for filterpy.kalman.KalmanFilter (third party, `requirements.txt:14` pins
        filterpy>=1.4.5; not installed here).  Published algorithm of 1.4.5: predict  x = F x, P = a^2 F P F^T + Q;
        update  y = z - H x, S = H P H^T + R, K = P H^T S^-1, x += K y, P = (I-KH) P (I-KH)^T + K R K^T (Joseph form).
        Only what utils/tracker.py uses: attributes x, P, Q, F, H, R and predict() / update(z, R=...)."""

        def __init__(self, dim_x, dim_z, dim_u=0):
            import numpy as _np
            self.dim_x, self.dim_z = dim_x, dim_z
            self.x = _np.zeros((dim_x, 1))
            self.P = _np.eye(dim_x)
            self.Q = _np.eye(dim_x)
            self.F = _np.eye(dim_x)
            self.H = _np.zeros((dim_z, dim_x))
            self.R = _np.eye(dim_z)
            self._alpha_sq = 1.0
            self._I = _np.eye(dim_x)

        def predict(self):
            import numpy as _np
            self.x = _np.dot(self.F, self.x)
            self.P = self._alpha_sq * _np.dot(_np.dot(self.F, self.P), self.F.T) + self.Q

        def update(self, z, R=None, H=None):
            import numpy as _np
            if z is None:
                return
            R = self.R if R is None else R
            H = self.H if H is None else H
            z = _np.atleast_2d(_np.asarray(z, dtype=float))
            if z.shape[1] == self.dim_z and z.shape[0] == 1:
                z = z.T
            y = z - _np.dot(H, self.x)
            PHT = _np.dot(self.P, H.T)
            S = _np.dot(H, PHT) + R
            K = _np.dot(PHT, _np.linalg.inv(S))
            self.x = self.x + _np.dot(K, y)
            I_KH = self._I - _np.dot(K, H)
            self.P = _np.dot(_np.dot(I_KH, self.P), I_KH.T) + _np.dot(_np.dot(K, R), K.T)
    fpk.KalmanFilter = KalmanFilter
    fpc.Q_discrete_white_noise = _not_impl
    fp.kalman = fpk
    fp.common = fpc
    sys.modules.setdefault("filterpy", fp)
    sys.modules.setdefault("filterpy.kalman", fpk)
    sys.modules.setdefault("filterpy.common", fpc)

    try:
        import sklearn.utils as sku
        la = types.ModuleType("sklearn.utils.linear_assignment_")
        def _linear_assignment(cost):
            # sklearn <= 0.22 API: array of (row, col) pairs of the optimal assignment, sorted by row
            import numpy as _np
            from scipy.optimize import linear_sum_assignment
            r, c = linear_sum_assignment(cost)
            return _np.stack([r, c], axis=1)
        la.linear_assignment = _linear_assignment
        sys.modules.setdefault("sklearn.utils.linear_assignment_", la)
        sku.linear_assignment_ = la
    except Exception:
        pass

    import torch.utils.model_zoo as model_zoo

    def _fake_load_url(url, *a, **k):
        return {"fc.weight": torch.zeros(1000, 512, 1, 1), "fc.bias": torch.zeros(1000)}
    model_zoo.load_url = _fake_load_url

    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None

    for p in (REF_LIB, REF_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    _installed = True


def make_opt(arch="dla_34", tracking_task=False, rep_mode=1, c="chair", gpus="-1",
             extra_args=()):
    """Build the reference `opt` exactly the way src/demo.py:93-155 does."""
    install()
    from lib.opts import opts
    argv = ["--arch", arch, "--gpus", gpus, "--rep_mode", str(rep_mode), "--c", c,
            "--debug", "0"] + list(extra_args)
    if tracking_task:
        argv.append("--tracking_task")
    saved = sys.argv
    sys.argv = ["demo.py"] + argv
    try:
        opt = opts().parser.parse_args()
    finally:
        sys.argv = saved
    opt.nms = True
    opt.obj_scale = True
    if opt.tracking_task:
        opt.pre_img = True
        opt.pre_hm = True
        opt.tracking = True
        opt.pre_hm_hp = True
        opt.tracking_hp = True
        opt.track_thresh = 0.1
        opt.obj_scale_uncertainty = True
        opt.hps_uncertainty = True
        opt.kalman = True
        opt.scale_pool = True
        opt.vis_thresh = max(opt.track_thresh, opt.vis_thresh)
        opt.pre_thresh = max(opt.track_thresh, opt.pre_thresh)
        opt.new_thresh = max(opt.track_thresh, opt.new_thresh)
    opt.use_pnp = True
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        opt = opts().parse(opt)
        opt = opts().init(opt)
    return opt


class legacy_bool_arith(object):
    """Context manager: torch 1.1.0 semantics for arithmetic on comparison results.

    The reference pins `torch==1.1.0` (requirements.txt:5, README.md:19).  There,
    `a > b` returns a uint8 tensor and `+` on those is an INTEGER add, so
    `mask_2 == 7` at models/decode.py:183-188 means "all seven gates hold".
    On torch >= 1.2 comparisons return bool and bool + bool is a logical OR,
    which silently turns `mask_2 == 7` into all-False (the heat-map keypoint
    representation is then never used).  The product and the oracle implement
    the pinned-version (intended) semantics; this shim makes the unmodified
    reference do the same on the torch in this image."""

    def __enter__(self):
        self._add = torch.Tensor.__add__
        orig = self._add

        def add(a, b):
            if isinstance(a, torch.Tensor) and a.dtype == torch.bool:
                a = a.to(torch.uint8)
            if isinstance(b, torch.Tensor) and b.dtype == torch.bool:
                b = b.to(torch.uint8)
            return orig(a, b)
        torch.Tensor.__add__ = add
        return self

    def __exit__(self, *exc):
        torch.Tensor.__add__ = self._add
        return False
