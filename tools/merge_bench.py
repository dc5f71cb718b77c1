"""Merge kernel timing at BASELINE size (169 x 512 x 512 float32 tiles -> 5000 x 5000 float32 + uint8 mask), periodic
kernels (TMA-staged with 4/2/1 pixels per thread, register-only) vs gather kernel (SNB_MERGE_GATHER=1), CUDA events, L2 flushed between iterations."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import snb_b200  # noqa: E402,F401
from snb_b200 import _native as N  # noqa: E402
from snb_b200.lib.tiles import ImageSlicer  # noqa: E402

lib, st = N.lib(), N.stream_ptr()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for (H, W, T, S) in [(5000, 5000, 512, 384), (5000, 5000, 224, 112)][:int(os.environ.get("MERGE_GEOMS", "2"))]:
    s = ImageSlicer((H, W, 1), T, S, weight="pyramid")
    n = len(s.crops)
    g = torch.Generator(device="cuda").manual_seed(0)
    probs = [torch.rand((n, T, T, 1), device="cuda", generator=g) for _ in range(2)]
    merged = torch.empty((H, W, 1), dtype=torch.float32, device="cuda")
    mask = torch.empty((H, W, 1), dtype=torch.uint8, device="cuda")
    wdev = s.weight_on_device()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    nbytes = n * T * T * 4 + H * W * 5
    for mode in os.environ.get("MERGE_MODES", "ring1,ring2,ring4,staged2,staged1,period,gather").split(","):
        os.environ.pop("SNB_MERGE_GATHER", None)
        os.environ["SNB_MERGE_MODE"] = mode
        if mode == "gather":
            os.environ["SNB_MERGE_GATHER"] = "1"
        best, tot = 1e9, 0.0
        for r in range(reps + 1):
            flush.fill_(r)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            N.check(lib.snb_merge(s.handle, N.ptr(probs[r % 2]), N.DT_F32, 1, 1, N.ptr(wdev), N.ptr(merged), N.DT_F32,
                                  N.ptr(mask), 0.5, st))
            e1.record()
            torch.cuda.synchronize()
            if r:
                ms = e0.elapsed_time(e1)
                best, tot = min(best, ms), tot + ms
        print("%d/%d %-7s avg %.3f ms best %.3f ms  %.1f MB  %.0f GB/s (avg)" % (T, S, mode, tot / reps, best, nbytes / 1e6,
                                                                                 nbytes / (tot / reps) / 1e6), flush=True)
