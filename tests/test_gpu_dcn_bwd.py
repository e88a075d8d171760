"""Row f-4 on the GPU: cp_dcn_v2_backward (through the `_ext.dcn_v2_backward` signature) vs the golden vectors of the
reference's own C++ backward, and at a CenterPose layer shape vs an fp64 autograd evaluation of the operator."""
import numpy as np
import pytest
import torch

import centerpose_b200 as cpb
from tests.test_oracle_dcn_bwd import CASES, KEYS, _inputs
from tests.util import golden

pytestmark = pytest.mark.gpu

# fp32 sums in another order (warp reductions, split position ranges, atomics): relative to the largest entry
TOL = {"fp32": 3e-6, "tf32x3": 3e-6}


def _run(x, off, mask, w, go, precision):
    dev = torch.device("cuda")
    t = [torch.from_numpy(v).to(dev) for v in (x, w, off, mask, go)]
    bias = torch.zeros(w.shape[0], device=dev)
    gi, goff, gm, gw, gb = cpb.dcn_v2_backward(t[0], t[1], bias, t[2], t[3], t[4], 3, 3, 1, 1, 1, 1, 1, 1, 1,
                                               precision=precision)
    torch.cuda.synchronize()
    return dict(zip(KEYS, (v.cpu().numpy() for v in (gi, goff, gm, gw, gb))))


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("name", CASES)
def test_matches_reference_golden(name, precision, cplib):
    g = golden(name)
    x, off, mask, w, go = _inputs(g)
    got = _run(x, off, mask, w, go, precision)
    for k in KEYS:
        scale = max(1.0, np.abs(g[k]).max())
        assert np.abs(got[k] - g[k]).max() <= TOL[precision] * scale * np.sqrt(x.shape[1]), (k, np.abs(got[k] - g[k]).max())


def test_layer_shape_vs_fp64_autograd(cplib):
    """64 -> 64 channels at 64 x 64, batch 3 (an IDA node of dla_34 at a quarter of the benched map), offsets of a few
    pixels: against torchvision's deform_conv2d differentiated in fp64 on the same device."""
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(5)
    B, C, H, W, Co = 3, 64, 64, 64, 64
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    off = (rng.standard_normal((B, 18, H, W)) * 2.0).astype(np.float32)
    mask = rng.random((B, 9, H, W)).astype(np.float32)
    w = (rng.standard_normal((Co, C, 3, 3)) / np.sqrt(C * 9)).astype(np.float32)
    go = rng.standard_normal((B, Co, H, W)).astype(np.float32)
    got = _run(x, off, mask, w, go, "fp32")
    t = [torch.from_numpy(v).cuda().double().requires_grad_(True) for v in (x, off, mask, w)]
    b = torch.zeros(Co, dtype=torch.float64, device="cuda", requires_grad=True)
    y = tv.ops.deform_conv2d(t[0], t[1], t[3], b, padding=1, mask=t[2])
    y.backward(torch.from_numpy(go).cuda().double())
    want = dict(zip(KEYS, (t[0].grad, t[1].grad, t[2].grad, t[3].grad, b.grad)))
    for k in KEYS:
        ref = want[k].cpu().numpy()
        err = np.abs(got[k] - ref).max() / max(1.0, np.abs(ref).max())
        assert err <= 2e-5, (k, err)
    # a second call gives the same deterministic parts bit for bit (grad_input uses atomics)
    again = _run(x, off, mask, w, go, "fp32")
    for k in ("grad_offset", "grad_mask", "grad_weight", "grad_bias"):
        assert np.array_equal(got[k], again[k]), k


def test_rejects_other_configurations(cplib):
    t = torch.zeros(1, 16, 8, 8, device="cuda")
    with pytest.raises(RuntimeError):
        cpb.dcn_v2_backward(t, torch.zeros(16, 16, 3, 3, device="cuda"), torch.zeros(16, device="cuda"),
                            torch.zeros(1, 18, 8, 8, device="cuda"), torch.zeros(1, 9, 8, 8, device="cuda"), t,
                            3, 3, 2, 2, 1, 1, 1, 1, 1)
