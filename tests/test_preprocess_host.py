"""CPU: the numpy restatement of OpenCV's fixed-point warpAffine (oracle/preprocess_ref.py) against cv2 itself -- the
pin for the bit-exact batched pre-process kernel."""
import numpy as np
import pytest

from oracle import preprocess_ref as pr

cv2 = pytest.importorskip("cv2")
SHAPES = [(600, 800), (480, 640), (512, 512), (720, 1280), (375, 500), (800, 600)]


@pytest.mark.parametrize("shape", SHAPES)
def test_warp_restatement_is_bit_exact(shape):
    h, w = shape
    img = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    M = pr.fix_res_affine(h, w, 512, 512)
    want = cv2.warpAffine(img, M, (512, 512), flags=cv2.INTER_LINEAR)
    assert np.array_equal(pr.warp_affine_u8(img, M, 512, 512), want)


def test_rotated_and_anisotropic_affine():
    img = np.random.default_rng(5).integers(0, 256, (300, 420, 3), dtype=np.uint8)
    for M in (cv2.getRotationMatrix2D((210, 150), 17.0, 0.8), np.array([[0.7, 0.1, 13.3], [-0.05, 1.2, -20.0]])):
        want = cv2.warpAffine(img, M, (256, 192), flags=cv2.INTER_LINEAR)
        assert np.array_equal(pr.warp_affine_u8(img, M, 256, 192), want)


def test_closed_form_affine_matches_the_reference_construction():
    """fix_res_affine == cv2.getAffineTransform on the reference's three float32 control points (to 1e-13; the
    fixed-point walk only sees differences on exact rounding ties)."""
    from centerpose_b200.detector import affine_from_center_scale
    for (h, w) in SHAPES:
        c = np.array([w / 2., h / 2.], np.float32)
        M = affine_from_center_scale(c, float(max(h, w)), 512, 512)
        assert np.abs(M - pr.fix_res_affine(h, w, 512, 512)).max() <= 1e-13


def test_pre_process_matches_reference_expression():
    img = np.random.default_rng(1).integers(0, 256, (480, 640, 3), dtype=np.uint8)
    mean, std = [0.408, 0.447, 0.470], [0.289, 0.274, 0.278]
    M = pr.fix_res_affine(480, 640, 512, 512)
    inp = cv2.warpAffine(cv2.resize(img, (640, 480)), M, (512, 512), flags=cv2.INTER_LINEAR)
    want = ((inp / 255. - np.array(mean, np.float32).reshape(1, 1, 3)) / np.array(std, np.float32).reshape(1, 1, 3)).astype(np.float32)
    got = pr.pre_process(img, 512, 512, mean, std)
    assert np.array_equal(got[0], want.transpose(2, 0, 1))
