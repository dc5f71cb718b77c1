"""Per-layer device times of the UNet16 plan (CUDA events around every op, warm, averaged).
Usage (GPU box): SNB_CONV_MODE=2 python tools/layer_times.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snb_b200  # noqa: E402,F401
from snb_b200 import synth  # noqa: E402
from snb_b200 import _native as N  # noqa: E402
from snb_b200.engine import ConvOp  # noqa: E402
from snb_b200.lib import models as M  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 13
T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
name = sys.argv[3] if len(sys.argv) > 3 else "unet16"
if name == "fcdensenet67":
    m = M.FCDenseNet67(n_classes=1)
    m.load_state_dict(synth.fcdensenet_state_dict(seed=0))
elif name == "zf_unet":
    m = M.ZF_UNET()
    m.load_state_dict(synth.zf_unet_state_dict(seed=0))
else:
    m = getattr(M, {"unet16": "UNet16", "unet11": "UNet11"}[name])()
    m.load_state_dict(synth.vgg_unet_state_dict(name, seed=0))
m = m.cuda().eval()
plan = m.plan(batch, T, T, sigmoid=True)
(plan.x_in3 if getattr(plan, 'x_in3', None) is not None else plan.x_patch).t.normal_()
reps = 5
for _ in range(2):
    plan.run()
torch.cuda.synchronize()
acc = [0.0] * len(plan.ops)
st = N.stream_ptr()
for _ in range(reps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(plan.ops) + 1)]
    ev[0].record()
    for k, op in enumerate(plan.ops):
        op(st)
        ev[k + 1].record()
    torch.cuda.synchronize()
    for k in range(len(plan.ops)):
        acc[k] += ev[k].elapsed_time(ev[k + 1]) / reps
tot = sum(acc)
print("mode=%s batch=%d tile=%d total %.3f ms  (%.1f TFLOP/s)" % (os.environ.get("SNB_CONV_MODE", "default"), batch, T, tot,
                                                                plan.flops / tot / 1e9))
for k, op in enumerate(plan.ops):
    if isinstance(op, ConvOp):
        d = op.desc
        print("%2d conv kind=%d %4dx%-4d cin=%4d cout=%4d  %8.3f ms %8.1f TF/s" % (
            k, d[0], d[1], d[2], d[3], d[4], acc[k], op.flops / acc[k] / 1e9))
    else:
        print("%2d %-8s %35s %8.3f ms" % (k, type(op).__name__, "", acc[k]))
by = {}
for k, op in enumerate(plan.ops):
    by[type(op).__name__] = by.get(type(op).__name__, 0.0) + acc[k]
print("by op type:", {k: round(v, 3) for k, v in by.items()})
