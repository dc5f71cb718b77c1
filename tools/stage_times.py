"""Device time of each pipeline stage of one 5000x5000 image (CUDA events, warm)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snb_b200  # noqa: E402,F401
from snb_b200 import synth  # noqa: E402
from snb_b200 import _native as N  # noqa: E402
from snb_b200 import inria_submit as sub  # noqa: E402
from snb_b200.lib import metrics  # noqa: E402
from snb_b200.lib.models import UNet16  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 13
m = UNet16()
m.load_state_dict(synth.vgg_unet_state_dict("unet16", seed=0))
m = m.cuda().eval()
pred = sub.TiledPredictor(m, (5000, 5000, 3), 512, 384, batch_size=batch, tta=False)
img = torch.from_numpy(synth.image_u8(0, 5000, 5000)).cuda()
gt = (torch.rand((5000, 5000, 1), device="cuda") > 0.5).to(torch.uint8)
lib, st = N.lib(), N.stream_ptr()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def split():
    layout, target = pred.plan.input_layout()
    N.check(lib.snb_split_norm_u8(pred.slicer.handle, N.ptr(img), 3, N.ptr(pred.lut), 0, layout, N.c_vp(target), 0, batch, st))


def copy():
    pred.probs[0:batch, 0, :, :, 0].copy_(pred.plan.out[:batch])


def merge():
    N.check(lib.snb_merge(pred.slicer.handle, N.ptr(pred.probs), N.DT_F32, 1, 1, N.ptr(pred.weight), N.ptr(pred.merged),
                          N.DT_F32, N.ptr(pred.mask), 0.5, st))


n_chunks = (169 + batch - 1) // batch
t_split, t_plan, t_copy, t_merge = timed(split), timed(pred.plan.run), timed(copy), timed(merge)
t_cnt = timed(lambda: metrics.confusion_counts_from_probs(pred.merged, gt))
t_all = timed(lambda: pred.predict_device(img), 3)
print("batch %d: split %.3f ms x%d, plan %.3f ms x%d, copy %.3f ms x%d, merge %.3f ms, counts %.3f ms" % (
    batch, t_split, n_chunks, t_plan, n_chunks, t_copy, n_chunks, t_merge, t_cnt))
print("sum of stages %.2f ms, predict_device %.2f ms -> %.1f Mpx/s" % (
    n_chunks * (t_split + t_plan + t_copy) + t_merge, t_all, 25.0 / t_all * 1e3))
