"""ctypes binding of libcenterpose_b200.so (include/centerpose_b200.h).

The product path has NO fallback: if the shared library is missing or a call
fails, a RuntimeError is raised.  (`python -m centerpose_b200.build` or
`__graft_entry__.build()` produces the library in-tree.)
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CP_LIB_PATH") or os.path.join(HERE, "libcenterpose_b200.so")      # CP_LIB_PATH: A/B another build

CP_MAX_HEADS = 16
CP_POSE_RECORD = 192
CP_DETS_RECORD = 128
CP_META_DOUBLES = 16
CP_MAX_K = 128

CP_ARCH_DLA34 = 0
CP_ARCH_DLAV1_34 = 1
CP_PREC_FP32 = 0
CP_PREC_TF32X3 = 1
CP_PREC_BF16 = 2
CP_PREC_TF32 = 3
PRECISIONS = {"fp32": CP_PREC_FP32, "tf32x3": CP_PREC_TF32X3, "bf16": CP_PREC_BF16, "tf32": CP_PREC_TF32}

# cp_pose_field
P_SCORE, P_CLS, P_STATUS, P_NPTS, P_BBOX, P_CT, P_KPS = 0, 1, 2, 3, 4, 8, 10
P_KPS_DISP_MEAN, P_KPS_HM_MEAN, P_KPS_HM_STD, P_KPS_HM_HEIGHT, P_KPS_DISP_STD = 26, 42, 58, 74, 82
P_OBJ_SCALE, P_OBJ_SCALE_UNC, P_TRACKING, P_TRACKING_HP = 98, 101, 104, 106
P_LOCATION, P_QUAT, P_REPROJ, P_PROJ_CUBOID, P_KPS_3D_CAM, P_KPS_PNP, P_SRC_INDEX = 122, 125, 129, 130, 146, 173, 191
# cp_dets_field
D_BBOX, D_SCORE, D_CLS, D_KPS, D_OBJ_SCALE, D_OBJ_SCALE_UNC, D_TRACKING, D_TRACKING_HP = 0, 4, 5, 6, 22, 25, 28, 30
D_KPS_DISP_MEAN, D_KPS_DISP_STD, D_KPS_HM_MEAN, D_KPS_HM_STD, D_KPS_HM_HEIGHT, D_IND = 46, 62, 78, 94, 110, 118
# cp_track_field
CP_TRACK_RECORD = 320
T_ID, T_AGE, T_ACTIVE, T_IN_BOXES, T_PNP2_STATUS, T_CONF_AVG = 192, 193, 194, 195, 196, 197
T_KPS_FUSION_MEAN, T_KPS_FUSION_STD, T_KPS_MEAN_KF, T_KPS_STD_KF = 200, 216, 232, 248
T_OBJ_SCALE_KF, T_OBJ_SCALE_UNC_KF, T_KPS_PNP_KF, T_KPS_3D_CAM_KF = 264, 267, 270, 288
# cp_pnp_status
PNP_NOT_RUN, PNP_OK, PNP_INVISIBLE, PNP_BEHIND, PNP_FEW_POINTS, PNP_SOLVER_FAIL = 0, 1, 2, 3, 4, 5

EXPORTS = [
    "cp_version", "cp_last_error", "cp_plan_create", "cp_plan_destroy", "cp_plan_load_weights",
    "cp_forward", "cp_plan_bytes", "cp_plan_forward_launches", "cp_decode_workspace_bytes",
    "cp_decode_pnp", "cp_infer", "cp_dcn_v2_forward", "cp_preprocess", "cp_plan_num_ops", "cp_plan_profile",
    "cp_dcn_v2_forward_ex", "cp_conv2d", "cp_dcn_v2_backward",
    "cp_preprocess_affine", "cp_tracker_create", "cp_tracker_destroy", "cp_tracker_reset", "cp_tracker_step", "cp_tracker_render",
]


class CpOpStat(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 96), ("kind", ctypes.c_int32), ("ms", ctypes.c_float),
                ("flops", ctypes.c_double), ("bytes", ctypes.c_double)]


class CpConfig(ctypes.Structure):
    _fields_ = [
        ("arch", ctypes.c_int32), ("tracking", ctypes.c_int32), ("tracking_task_gru", ctypes.c_int32),
        ("max_batch", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
        ("precision", ctypes.c_int32), ("device", ctypes.c_int32), ("head_conv", ctypes.c_int32),
        ("num_heads", ctypes.c_int32),
        ("head_names", ctypes.c_char_p * CP_MAX_HEADS),
        ("head_channels", ctypes.c_int32 * CP_MAX_HEADS),
    ]


class CpHeads(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "hm", "wh", "hps", "reg", "hm_hp", "hp_offset", "scale", "hps_uncertainty",
        "scale_uncertainty", "tracking", "tracking_hp")]


class CpDecodeParams(ctypes.Structure):
    _fields_ = [
        ("batch", ctypes.c_int32), ("out_h", ctypes.c_int32), ("out_w", ctypes.c_int32),
        ("num_classes", ctypes.c_int32), ("num_joints", ctypes.c_int32), ("K", ctypes.c_int32),
        ("rep_mode", ctypes.c_int32), ("use_moments", ctypes.c_int32), ("nms", ctypes.c_int32),
        ("visible_thresh", ctypes.c_int32), ("opencv_return", ctypes.c_int32),
        ("apply_sigmoid", ctypes.c_int32), ("use_pnp", ctypes.c_int32),
        ("vis_thresh", ctypes.c_float), ("balance", ctypes.c_float), ("modern_bool_semantics", ctypes.c_int32),
        ("test_scale", ctypes.c_float), ("num_scales", ctypes.c_int32),
    ]


class CpTrackerConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "streams", "max_tracks", "kalman", "scale_pool", "use_pnp", "hps_uncertainty", "max_age", "visible_thresh",
        "opencv_return", "render_hm_mode", "render_hmhp_mode", "device")] + [
        (n, ctypes.c_float) for n in ("new_thresh", "pre_thresh", "R", "conf_lo", "conf_hi")]


_lib = None


def lib_available():
    return os.path.exists(LIB_PATH)


def load():
    """Load the shared library (once).  Raises RuntimeError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "centerpose_b200: %s not found -- run `python -m centerpose_b200.build` "
            "(there is no CPU / PyTorch fallback for the hot path)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    L.cp_version.restype = ctypes.c_int
    L.cp_last_error.restype = ctypes.c_char_p
    L.cp_plan_create.argtypes = [ctypes.POINTER(CpConfig), ctypes.POINTER(vp)]
    L.cp_plan_destroy.argtypes = [vp]
    L.cp_plan_load_weights.argtypes = [vp, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(vp),
                                       ctypes.POINTER(i64), i32, vp]
    L.cp_forward.argtypes = [vp, i32, vp, vp, vp, vp, ctypes.POINTER(vp), vp]
    L.cp_plan_num_ops.argtypes = [vp]
    L.cp_plan_profile.argtypes = [vp, i32, vp, vp, vp, vp, ctypes.POINTER(vp), vp, ctypes.POINTER(CpOpStat), i32,
                                  ctypes.POINTER(i32)]
    L.cp_plan_bytes.argtypes = [vp]
    L.cp_plan_bytes.restype = i64
    L.cp_plan_forward_launches.argtypes = [vp]
    L.cp_plan_forward_launches.restype = i32
    L.cp_decode_workspace_bytes.argtypes = [ctypes.POINTER(CpDecodeParams)]
    L.cp_decode_workspace_bytes.restype = ctypes.c_size_t
    L.cp_decode_pnp.argtypes = [ctypes.POINTER(CpDecodeParams), ctypes.POINTER(CpHeads), vp, vp, vp, vp, vp,
                                ctypes.c_size_t, vp]
    L.cp_infer.argtypes = [vp, i32, vp, vp, vp, vp, ctypes.POINTER(CpDecodeParams), vp,
                           ctypes.POINTER(vp), vp, vp, vp, vp]
    L.cp_dcn_v2_forward.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    L.cp_dcn_v2_forward_ex.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    L.cp_dcn_v2_backward.argtypes = [vp] * 10 + [i32] * 6 + [vp]
    L.cp_conv2d.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]
    L.cp_preprocess.argtypes = [vp, vp, i32, i32, i32, i32, i32, ctypes.POINTER(ctypes.c_float),
                                ctypes.POINTER(ctypes.c_float), vp]
    L.cp_preprocess_affine.argtypes = [vp, vp, i32, i32, i32, i32, i32, ctypes.POINTER(ctypes.c_double),
                                       ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), vp]
    L.cp_tracker_create.argtypes = [ctypes.POINTER(CpTrackerConfig), ctypes.POINTER(vp)]
    L.cp_tracker_destroy.argtypes = [vp]
    L.cp_tracker_reset.argtypes = [vp, i32, vp]
    L.cp_tracker_step.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp, vp]
    L.cp_tracker_render.argtypes = [vp, i32, vp, vp, i32, i32, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("cp_version", "cp_last_error", "cp_plan_bytes", "cp_plan_forward_launches",
                        "cp_decode_workspace_bytes"):
            fn.restype = ctypes.c_int
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = load().cp_last_error()
        raise RuntimeError("centerpose_b200 %s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
