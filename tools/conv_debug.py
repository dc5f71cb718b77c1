"""Diagnostic runner for the tcgen05 convolution: prints error structure instead of a bare assert.
Usage (GPU box): python tools/conv_debug.py"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snb_b200  # noqa: E402,F401
from snb_b200 import _native as N  # noqa: E402
from snb_b200 import engine as E  # noqa: E402


def bf(t):
    return t.to(torch.bfloat16).float()


def run(kind, n, h, w, cin, cout, relu=True):
    g = torch.Generator(device="cuda").manual_seed(1)
    src = E.Slab(n, h, w, cin, "cuda")
    src.t.copy_(torch.randn((n, h, w, cin), device="cuda", generator=g).to(torch.bfloat16))
    s = 2 if kind == N.CONVT_4X4_S2 else 1
    dst = E.Slab(n, h * s, w * s, cout, "cuda")
    dst.t.fill_(float("nan"))
    bias = torch.randn(cout, device="cuda", generator=g)
    x = src.t.float().permute(0, 3, 1, 2)
    if kind == N.CONV_3X3:
        wt = torch.randn((cout, cin, 3, 3), device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5
        packed = E.pack_conv3x3(wt)
        want = F.conv2d(x, bf(wt), bias, padding=1)
    elif kind == N.CONV_1X1:
        wt = torch.randn((cout, cin, 1, 1), device="cuda", generator=g) * (2.0 / cin) ** 0.5
        packed = E.pack_conv1x1(wt)
        want = F.conv2d(x, bf(wt), bias)
    else:
        wt = torch.randn((cin, cout, 4, 4), device="cuda", generator=g) * (2.0 / (4 * cin)) ** 0.5
        packed = E.pack_convT4x4(wt)
        want = F.conv_transpose2d(x, bf(wt), bias, stride=2, padding=1)
    if relu:
        want = F.relu(want)
    op = E.ConvOp(kind, src.view(), dst.view(), packed, bias, relu=relu)
    op(N.stream_ptr())
    torch.cuda.synchronize()
    got = dst.t.float().permute(0, 3, 1, 2)
    nan = torch.isnan(got).sum().item()
    err = (got - want).abs()
    err = torch.where(torch.isnan(err), torch.full_like(err, 1e9), err)
    scale = want.abs().max().item()
    bad = err > 2e-2 * scale
    print("kind=%d n=%d h=%d w=%d cin=%d cout=%d: max_err=%.4g scale=%.4g nan=%d bad=%d/%d" % (
        kind, n, h, w, cin, cout, err.max().item(), scale, nan, bad.sum().item(), bad.numel()), flush=True)
    if bad.any():
        per_c = bad.sum(dim=(0, 2, 3))
        per_y = bad.sum(dim=(0, 1, 3))
        per_x = bad.sum(dim=(0, 1, 2))
        print("  bad per channel (first 64):", per_c[:64].tolist())
        print("  bad per y:", per_y.tolist()[:64])
        print("  bad per x:", per_x.tolist()[:64])
        print("  got[0,:4,0,:8]:", got[0, :4, 0, :8].tolist())
        print("  want[0,:4,0,:8]:", want[0, :4, 0, :8].tolist())
    return not bad.any()


if __name__ == "__main__":
    ok = True
    ok &= run(N.CONV_1X1, 1, 8, 16, 64, 64, relu=False)
    ok &= run(N.CONV_1X1, 1, 8, 16, 32, 32, relu=False)
    ok &= run(N.CONV_1X1, 1, 8, 16, 128, 256, relu=False)
    ok &= run(N.CONV_3X3, 1, 8, 16, 64, 64, relu=False)
    ok &= run(N.CONV_3X3, 2, 32, 32, 64, 128)
    ok &= run(N.CONV_3X3, 1, 24, 40, 256, 512)
    ok &= run(N.CONV_3X3, 3, 16, 32, 96, 32)
    ok &= run(N.CONVT_4X4_S2, 1, 8, 16, 64, 64)
    ok &= run(N.CONVT_4X4_S2, 2, 16, 16, 512, 256)
    ok &= run(N.CONV_3X3, 4, 128, 128, 256, 256)
    print("ALL OK" if ok else "FAILURES")
    sys.exit(0 if ok else 1)
