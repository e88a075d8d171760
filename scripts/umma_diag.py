"""Layout diagnostics for the tcgen05 kernel: structured operands whose product exposes any
descriptor / swizzle / TMEM-lane mistake as a visible permutation.  Writes gpurun_out/umma_diag.npz."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import centerpose_b200 as cpb  # noqa: E402

out = {}
for prec in ("bf16", "tf32x3"):
    for (Cin, Cout, M) in ((64, 64, 128), (128, 32, 256), (64, 256, 128)):
        H, W = M // 16, 16
        # A = small integers exactly representable in bf16; W = one-hot rows -> out[m, n] = A[m, sel(n)]
        a = torch.arange(M * Cin, dtype=torch.float32).reshape(1, H, W, Cin) % 251 - 125
        sel = (torch.arange(Cout) * 7 + 3) % Cin
        w = torch.zeros(Cout, Cin, 1, 1)
        w[torch.arange(Cout), sel, 0, 0] = 1.0
        try:
            got = cpb.conv2d_nhwc(a.cuda(), w.cuda(), None, None, precision=prec)
            torch.cuda.synchronize()
            got = got.cpu().reshape(M, Cout)
            want = a.reshape(M, Cin)[:, sel]
            bad = (got != want)
            print("%s Cin=%d Cout=%d M=%d: mismatches %d / %d" % (prec, Cin, Cout, M, int(bad.sum()), bad.numel()))
            if bad.any():
                idx = bad.nonzero()[:8]
                for i, j in idx.tolist():
                    print("   out[%d,%d] = %g want %g" % (i, j, got[i, j].item(), want[i, j].item()))
            out["%s_%d_%d_%d_got" % (prec, Cin, Cout, M)] = got.numpy()
            out["%s_%d_%d_%d_want" % (prec, Cin, Cout, M)] = want.numpy()
        except Exception as e:   # noqa: BLE001
            print("%s Cin=%d Cout=%d M=%d: FAILED %s" % (prec, Cin, Cout, M, e))
            break
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/umma_diag.npz", **out)
