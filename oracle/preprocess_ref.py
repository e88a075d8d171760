"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's pre-process
(/root/reference/src/lib/detectors/base_detector.py:91-148): cv2.resize to the same size (identity copy),
cv2.warpAffine(INTER_LINEAR) with `trans_input`, then ((x / 255.) - mean) / std -> float32 NCHW.

The warp arithmetic lives in OpenCV (third party, not under /root/reference: `opencv-python>=4.5.3.56`,
requirements.txt:11; 4.13.0 in this image).  Its published fixed-point algorithm (imgwarp.cpp: WarpAffineInvoker +
remapBilinear for CV_8U) is restated in `warp_affine_u8`:
  inverse matrix in double; X0 = cvRound((M[1] y + M[2]) * 1024) + 16; adelta[x] = cvRound(M[0] x * 1024);
  X = (X0 + adelta[x]) >> 5; source pixel X >> 5, fraction X & 31 (1/32 px); integer weights (32-fy)(32-fx)*32 ...
  (sum 2^15); value = (sum w * src + 2^14) >> 15; BORDER_CONSTANT 0.
Parity status: PINNED -- tests/test_preprocess_host.py checks it bit for bit against cv2.warpAffine itself on the
Objectron frame shapes (600x800, 480x640, 512x512, ...).
"""
import numpy as np


def invert_affine(M):
    """cv::warpAffine's own inversion of the forward 2x3 matrix (same operation order)."""
    M = np.asarray(M, np.float64).reshape(-1).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    return M


def warp_affine_u8(src, M, dw, dh):
    """uint8 [H,W,C] -> uint8 [dh,dw,C]: cv2.warpAffine(src, M, (dw, dh), flags=cv2.INTER_LINEAR)."""
    Mi = invert_affine(M)
    rnd = lambda v: np.rint(v).astype(np.int64)          # noqa: E731  cvRound: round half to even
    xs = np.arange(dw, dtype=np.float64)
    adelta, bdelta = rnd(Mi[0] * xs * 1024.0), rnd(Mi[3] * xs * 1024.0)
    H, W = src.shape[:2]
    S = src.astype(np.int64)
    out = np.zeros((dh, dw, src.shape[2]), np.uint8)
    for y in range(dh):
        X0 = int(rnd((Mi[1] * y + Mi[2]) * 1024.0)) + 16
        Y0 = int(rnd((Mi[4] * y + Mi[5]) * 1024.0)) + 16
        X, Y = (X0 + adelta) >> 5, (Y0 + bdelta) >> 5
        sx, sy, fx, fy = X >> 5, Y >> 5, X & 31, Y & 31
        w = [(32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32]

        def px(yy, xx):
            ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            return S[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)] * ok[:, None]
        t = (px(sy, sx) * w[0][:, None] + px(sy, sx + 1) * w[1][:, None] + px(sy + 1, sx) * w[2][:, None] +
             px(sy + 1, sx + 1) * w[3][:, None])
        out[y] = np.clip((t + (1 << 14)) >> 15, 0, 255).astype(np.uint8)
    return out


def fix_res_affine(height, width, inp_w, inp_h):
    """trans_input of the fix_res branch (base_detector.py:109-121) in closed form (utils/image.py:35-68, rot = 0)."""
    c = np.array([width / 2., height / 2.], np.float32)
    s = np.float32(max(height, width) * 1.0)
    src1y = np.float32(c[1] + s * np.float32(-0.5))
    dst0 = np.array([inp_w * 0.5, inp_h * 0.5], np.float32)
    dst1y = np.float32(dst0[1] + np.float32(inp_w * -0.5))
    a = (float(dst1y) - float(dst0[1])) / (float(src1y) - float(c[1]))
    return np.array([[a, 0.0, float(dst0[0]) - a * float(c[0])], [0.0, a, float(dst0[1]) - a * float(c[1])]])


def pre_process(image, inp_w, inp_h, mean, std, trans_input=None):
    """-> float32 [1,3,inp_h,inp_w] like base_detector.py:128-136."""
    h, w = image.shape[:2]
    M = fix_res_affine(h, w, inp_w, inp_h) if trans_input is None else np.asarray(trans_input, np.float64)
    inp = warp_affine_u8(image, M, inp_w, inp_h)
    mean = np.asarray(mean, np.float32).reshape(1, 1, 3)
    std = np.asarray(std, np.float32).reshape(1, 1, 3)
    x = ((inp / 255. - mean) / std).astype(np.float32)
    return x.transpose(2, 0, 1)[None]
