"""One conv layer through cp_conv2d for ncu:  CP_SHAPE=B,H,W,Cin,Cout,k  CP_PRECISION=tf32x3 python scripts/profile_conv.py
(two launches: the first warms up, profile the second:  ncu --set full -k regex:conv_tma -s 1 -c 1 ...)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import centerpose_b200 as cpb  # noqa: E402

B, H, W, Cin, Cout, k = [int(v) for v in os.environ.get("CP_SHAPE", "32,128,128,64,27,3").split(",")]
prec = os.environ.get("CP_PRECISION", "tf32x3")
g = torch.Generator().manual_seed(0)
x = torch.randn(B, H, W, Cin, generator=g).cuda()
w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).cuda()
b = torch.randn(Cout, generator=g).cuda()
for _ in range(2):
    y = cpb.conv2d_nhwc(x, w, b, None, stride=1, pad=k // 2, relu=True, precision=prec)
torch.cuda.synchronize()
print(float(y.abs().mean()))
