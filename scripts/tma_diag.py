"""Addressing diagnostics for the TMA shifted-window conv (conv_tma.cu): one-hot weights that pick a single
(tap, channel) make every output equal to one input element, so any slab / descriptor / base-offset mistake shows
up as an exact, locatable mismatch.  (This is the test that showed the descriptor base-offset field must stay 0: UMMA
swizzles on absolute shared-memory address bits)"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import centerpose_b200 as cpb  # noqa: E402

torch.manual_seed(0)
for bo in ("tf32", "tf32x3"):
    for (B, H, W, Cin, Cout, k) in ((1, 8, 16, 32, 32, 1), (1, 16, 16, 16, 16, 3), (2, 64, 64, 16, 16, 3), (1, 8, 16, 16, 32, 1), (1, 16, 16, 32, 64, 3), (2, 12, 30, 64, 32, 3),
                                    (1, 128, 128, 64, 256, 3), (1, 16, 16, 512, 128, 3)):
        x = (torch.arange(B * H * W * Cin, dtype=torch.float32).reshape(B, H, W, Cin) * 7 % 251) - 125
        w = torch.zeros(Cout, Cin, k, k)
        for n in range(Cout):
            w[n, (n * 5 + 1) % Cin, (n // 3) % k, n % k] = 1.0
        want = F.conv2d(x.permute(0, 3, 1, 2), w, None, 1, k // 2).permute(0, 2, 3, 1)
        try:
            got = cpb.conv2d_nhwc(x.cuda(), w.cuda(), None, None, stride=1, pad=k // 2, precision=bo)
            torch.cuda.synchronize()
            got = got.cpu()
            bad = got != want
            print("%s B%d %dx%d Cin%d Cout%d k%d: mismatches %d / %d" % (bo, B, H, W, Cin, Cout, k, int(bad.sum()),
                                                                                      bad.numel()))
            if bad.any():
                idx = bad.nonzero()[:6]
                for b, y, xx, n in idx.tolist():
                    print("    out[b%d y%d x%d n%d] = %g want %g" % (b, y, xx, n, got[b, y, xx, n].item(), want[b, y, xx, n].item()))
                rows_bad = bad.any(dim=3).sum().item()
                print("    positions with any error: %d of %d" % (rows_bad, B * H * W))
        except Exception as e:   # noqa: BLE001
            print("%s B%d %dx%d Cin%d Cout%d k%d: FAILED %s" % (bo, B, H, W, Cin, Cout, k, str(e)[:200]))
            sys.exit(0)          # a trap kills the context; stop here
# random-data accuracy (tf32 single pass): expect ~1e-3 of max
for (B, H, W, Cin, Cout, k) in ((2, 32, 32, 64, 64, 3), (2, 64, 64, 16, 16, 3), (1, 64, 64, 128, 256, 1), (4, 128, 128, 64, 1792, 3), (1, 8, 8, 512, 256, 3)):
    x = torch.randn(B, H, W, Cin)
    w = torch.randn(Cout, Cin, k, k) / np.sqrt(Cin * k * k)
    b = torch.randn(Cout)
    want = F.relu(F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), 1, k // 2)).permute(0, 2, 3, 1).float()
    for prec in ("tf32", "tf32x3"):
        got = cpb.conv2d_nhwc(x.cuda(), w.cuda(), b.cuda(), None, stride=1, pad=k // 2, relu=True, precision=prec).cpu()
        print("random %s B%d %dx%d Cin%d Cout%d k%d: rel err %.3e" % (prec, B, H, W, Cin, Cout, k, (got - want).abs().max().item() / want.abs().max().item()))
