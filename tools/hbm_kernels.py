"""Achieved HBM bandwidth of the byte/integer kernels at BASELINE sizes (5000x5000, 169 x 512 x 512).
CUDA events, warm, inputs far larger than L2 or rotated; algorithmic bytes per DESIGN.md section 3."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import snb_b200  # noqa: E402,F401
from snb_b200 import _native as N  # noqa: E402
from snb_b200.lib import losses, metrics  # noqa: E402
from snb_b200.lib.augmentations import NormalizeImage  # noqa: E402
from snb_b200.lib.tiles import ImageSlicer  # noqa: E402

peak = 6549.1
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
lib, st = N.lib(), N.stream_ptr()
H = W = 5000
T, S = 512, 384
s3 = ImageSlicer((H, W, 3), T, S, weight="pyramid")
n = len(s3.crops)
g = torch.Generator(device="cuda").manual_seed(0)
imgs = [torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(3)]
lut = torch.from_numpy(NormalizeImage(mean=[0.4, 0.45, 0.43], std=[3.1, 3.3, 3.6]).lut()).cuda()
rows = torch.empty((n, T, T, 32), dtype=torch.bfloat16, device="cuda")
tiles_u8 = torch.empty((n, T, T, 3), dtype=torch.uint8, device="cuda")
probs = [torch.rand((n, T, T, 1), device="cuda", generator=g) for _ in range(2)]
merged = torch.empty((H, W, 1), dtype=torch.float32, device="cuda")
mask = torch.empty((H, W, 1), dtype=torch.uint8, device="cuda")
wdev = s3.weight_on_device()
ne = n * T * T
logits = [torch.randn(ne, device="cuda", generator=g) for _ in range(2)]
t64 = [(torch.rand(ne, device="cuda", generator=g) > 0.5).long() for _ in range(2)]
t8 = [t.to(torch.uint8) for t in t64]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timed(fn, reps=6):
    fn(0)
    torch.cuda.synchronize()
    tot = 0.0
    for r in range(reps):
        flush.fill_(r)                         # L2 flush between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(r)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print("%-34s %8.3f ms  %8.1f MB  %8.1f GB/s  %5.1f%% of measured %.0f GB/s" % (name, ms, nbytes / 1e6, gbs,
                                                                                   100 * gbs / peak, peak), flush=True)


report("split_norm_u8 -> PATCH32", timed(lambda r: N.check(lib.snb_split_norm_u8(
    s3.handle, N.ptr(imgs[r % 3]), 3, N.ptr(lut), 0, N.LAYOUT_PATCH32, N.ptr(rows), 0, n, st))), H * W * 3 + n * T * T * 64)
report("split_hwc u8x3 (ImageSlicer.split)", timed(lambda r: N.check(lib.snb_split_hwc(
    s3.handle, N.ptr(imgs[r % 3]), 3, 1, 0, None, N.ptr(tiles_u8), 0, n, st))), H * W * 3 + n * T * T * 3)
report("merge f32 tiles -> f32 + u8 mask", timed(lambda r: N.check(lib.snb_merge(
    s3.handle, N.ptr(probs[r % 2]), N.DT_F32, 1, 1, N.ptr(wdev), N.ptr(merged), N.DT_F32, N.ptr(mask), 0.5, st))),
    n * T * T * 4 + H * W * 5)
report("loss_iou_reduce (f32 + int64)", timed(lambda r: losses.fused_sums(logits[r % 2], t64[r % 2])), ne * 12)
report("loss_iou_reduce (f32 + u8)", timed(lambda r: losses.fused_sums(logits[r % 2], t8[r % 2])), ne * 5)
report("confusion_counts (f32 + u8)", timed(lambda r: metrics.confusion_counts_from_probs(logits[r % 2], t8[r % 2])), ne * 5)
m = metrics.PRCurveMeter()
report("pr_curve_update (f32 + int64)", timed(lambda r: m.update(logits[r % 2], t64[r % 2])), ne * 12)
x8 = torch.randn(401408, device="cuda")
y8 = (torch.rand(401408, device="cuda") > 0.5).long()
report("loss_iou_reduce config-1 size", timed(lambda r: losses.fused_sums(x8, y8)), 401408 * 12)
