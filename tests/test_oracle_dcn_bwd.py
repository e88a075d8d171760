"""Row f-4 oracle: the golden vectors of the reference's own C++ CPU backward (oracle/_ref, dcn_v2_cpu.cpp:109-224 +
dcn_v2_im2col_cpu.cpp col2im / col2im_coord / im2col) are pinned against an independent fp64 autograd evaluation of the
same operator (torchvision.ops.deform_conv2d) and, in the build container, against the live oracle/_ref library."""
import os

import numpy as np
import pytest
import torch

from tests.util import ROOT, golden

CASES = ["dcn_bwd_small", "dcn_bwd_edge_big_offsets", "dcn_bwd_64ch"]
KEYS = ("grad_input", "grad_offset", "grad_mask", "grad_weight", "grad_bias")


def _inputs(g):
    from oracle.make_golden import dcn_bwd_inputs
    return dcn_bwd_inputs(*[g[k].item() for k in ("B", "C", "H", "W", "Co", "seed", "off_std")])


@pytest.mark.parametrize("name", CASES)
def test_golden_matches_fp64_autograd(name):
    tv = pytest.importorskip("torchvision")
    g = golden(name)
    x, off, mask, w, go = _inputs(g)
    t = [torch.from_numpy(v).double().requires_grad_(True) for v in (x, off, mask, w)]
    b = torch.zeros(w.shape[0], dtype=torch.float64, requires_grad=True)
    y = tv.ops.deform_conv2d(t[0], t[1], t[3], b, padding=1, mask=t[2])
    y.backward(torch.from_numpy(go).double())
    want = dict(zip(KEYS, (t[0].grad, t[1].grad, t[2].grad, t[3].grad, b.grad)))
    for k in KEYS:
        ref = want[k].numpy()
        assert np.abs(g[k] - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()) * np.sqrt(x.shape[1]), k


def test_golden_matches_live_reference_library():
    so = os.path.join(ROOT, "oracle", "_ref", "libdcn_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (GPU box / no reference tree)")
    from oracle.make_golden import ref_dcn_backward
    for name in CASES:
        g = golden(name)
        out = ref_dcn_backward(*_inputs(g))
        for k in KEYS:
            assert np.array_equal(out[k], g[k]), (name, k)
