"""Achieved HBM bandwidth of the byte / integer kernels of the hot path at BASELINE sizes (one 5000 x 5000 image = 169 tiles
of 512 x 512), for bench.py's `roofline_hbm` block and tools/hbm_kernels.py.

Timing: CUDA events on the launching stream around `reps` back-to-back calls after a warm-up call; every kernel moves more
than the 126 MB L2 per call and the inputs ROTATE over several sets (at least 2 x L2 in total), so a call never finds its
inputs in L2 ("inputs larger than L2" of the timing rules).  An explicit fill between the calls is deliberately NOT used:
it leaves 126 MB of dirty lines whose write-back is then charged to the next kernel (measured: +20-50 % on these 50-130 us
kernels).  `bytes` are the ALGORITHMIC bytes of SURVEY.md 8(d) (every input element read once, every output element
written once), `frac` = bytes / time / the measured copy bandwidth of MEASURED_PEAKS.json.
"""
import torch

from . import _native as N
from .lib import losses, metrics
from .lib.augmentations import NormalizeImage
from .lib.tiles import ImageSlicer

H = W = 5000
T, S = 512, 384


def measure(peak_gbs, reps=5, device=None, only=None):
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    lib, st = N.lib(), N.stream_ptr()
    s3 = ImageSlicer((H, W, 3), T, S, weight="pyramid")
    n = len(s3.crops)
    g = torch.Generator(device=dev).manual_seed(0)
    imgs = [torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, device=dev, generator=g) for _ in range(4)]   # 4 x 75 MB
    lut = torch.from_numpy(NormalizeImage(mean=[0.4, 0.45, 0.43], std=[3.1, 3.3, 3.6]).lut()).to(dev)
    ne = n * T * T
    out = {}

    def timed(fn):
        for r in range(2):
            fn(r)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            fn(r)
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    def report(name, ms, nbytes, what):
        gbs = nbytes / ms / 1e6
        out[name] = {"us": ms * 1e3, "bytes": nbytes, "GBps": gbs, "frac": gbs / peak_gbs, "what": what}

    def want(name):
        return only is None or name in only

    if want("split_norm"):
        dst = torch.empty((n, T, T, 3), dtype=torch.bfloat16, device=dev)
        layout = N.LAYOUT_NHWC3_BF16
        report("split_norm", timed(lambda r: N.check(lib.snb_split_norm_u8(
            s3.handle, N.ptr(imgs[r % 4]), 3, N.ptr(lut), 0, layout, N.ptr(dst), 0, n, st))),
            H * W * 3 + n * T * T * 3 * 2,
            "u8 image -> normalise -> reflect-101 pad -> 169 network-input tiles; SURVEY 8d: 75.0 MB read + 265.8 MB "
            "(3 bf16 channels per tile pixel) written")
        del dst
    if want("split_hwc"):
        tiles_u8 = torch.empty((n, T, T, 3), dtype=torch.uint8, device=dev)
        report("split_hwc", timed(lambda r: N.check(lib.snb_split_hwc(
            s3.handle, N.ptr(imgs[r % 4]), 3, 1, 0, None, N.ptr(tiles_u8), 0, n, st))), H * W * 3 + n * T * T * 3,
            "ImageSlicer.split of the u8 image (lib/tiles.py:98-117): 75 MB read + 132.9 MB written")
        del tiles_u8
    if want("merge"):
        probs = [torch.rand((n, T, T, 1), device=dev, generator=g) for _ in range(3)]
        merged = torch.empty((H, W, 1), dtype=torch.float32, device=dev)
        mask = torch.empty((H, W, 1), dtype=torch.uint8, device=dev)
        wdev = s3.weight_on_device()
        report("merge", timed(lambda r: N.check(lib.snb_merge(
            s3.handle, N.ptr(probs[r % 3]), N.DT_F32, 1, 1, N.ptr(wdev), N.ptr(merged), N.DT_F32, N.ptr(mask), 0.5, st))),
            n * T * T * 4 + H * W * 5, "ImageSlicer.merge + threshold: 177.2 MB of f32 tiles read, 100 MB f32 + 25 MB u8 written")
        del probs, merged, mask
    logits = [torch.randn(ne, device=dev, generator=g) for _ in range(3)]
    t8 = [(torch.rand(ne, device=dev, generator=g) > 0.5).to(torch.uint8) for _ in range(3)]
    if want("loss_i64") or want("pr_curve"):
        t64 = [t.long() for t in t8]
    if want("loss_i64"):
        report("loss_i64", timed(lambda r: losses.fused_sums(logits[r % 3], t64[r % 3])), ne * 12,
               "fused BCE + soft-Jaccard sums + IoU counts, f32 logits + int64 targets (the reference's dtype), one launch")
    if want("loss_u8"):
        report("loss_u8", timed(lambda r: losses.fused_sums(logits[r % 3], t8[r % 3])), ne * 5, "same with uint8 targets")
    if want("counts"):
        report("counts", timed(lambda r: metrics.confusion_counts_from_probs(logits[r % 3], t8[r % 3])), ne * 5,
               "confusion counts of f32 probabilities against u8 targets")
    if want("pr_curve"):
        m = metrics.PRCurveMeter()
        report("pr_curve", timed(lambda r: m.update(logits[r % 3], t64[r % 3])), ne * 12,
               "PRCurveMeter.update, 127 thresholds, f32 logits + int64 targets, one launch")
    return out
