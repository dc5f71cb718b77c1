"""Where the roles of conv_halo_kernel wait, per UNet16 layer (needs the profiling build):
    python tools/build_rev.py --profile
    SNB_B200_LIB=tools/ab/libsnb_b200_profile.so python tools/conv_wait_profile.py [batch]
Prints, per layer, cycles per tile: issuer loop, issuer waits (activation stage / weight slot / accumulator stage), epilogue
warp loop and its wait for a full accumulator, weight producer's wait for a free slot."""
import ctypes
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
else:
    m = M.UNet16()
    m.load_state_dict(synth.vgg_unet_state_dict("unet16", seed=0))
m = m.cuda().eval()
plan = m.plan(batch, T, T, sigmoid=True)
(plan.x_in3 if getattr(plan, 'x_in3', None) is not None else plan.x_patch).t.normal_()
for _ in range(2):
    plan.run()
torch.cuda.synchronize()
lib = N.lib()
prof = lib.snb_debug_conv_profile
prof.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
prof_sc = lib.snb_debug_scatter_profile      # conv_scatter_kernel: "w:wgt" = prologue warp waiting for the raw stage, "prodB" = prologue loop
prof_sc.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
buf = (ctypes.c_ulonglong * 12)()
st = N.stream_ptr()
sms = torch.cuda.get_device_properties(0).multi_processor_count
print("cycles per tile (one CTA's view; %d CTAs)" % sms)
print("%-44s %7s | %8s %7s %7s %7s | %8s %7s | %7s" % ("layer", "tiles", "issuer", "w:act", "w:wgt", "w:acc", "epilogue", "w:full", "prodB"))
for k, op in enumerate(plan.ops):
    if not getattr(op, "flops", 0):
        continue
    torch.cuda.synchronize()
    prof(buf, 1)
    prof_sc(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    op(st)
    e1.record()
    torch.cuda.synchronize()
    prof(buf, 0)
    v = list(buf)
    kind = "halo"
    if v[7] == 0:
        prof_sc(buf, 0)
        v = list(buf)
        kind = "scat"
    tiles = max(1, v[7])
    d = getattr(op, "desc", None)
    name = ("%2d %s k=%d %4dx%-4d %4d->%-4d" % (k, kind, d[0], d[1], d[2], d[3], d[4]) if d is not None else "%2d %s %s" % (k, kind, type(op).__name__)) + " %.3f ms" % e0.elapsed_time(e1)
    extra = "  | TS prologue: w:tmem %.0f  st+arrive %.0f" % (v[8] / tiles, v[9] / tiles) if kind == "scat" and v[9] else ""
    print("%-44s %7d | %8.0f %7.0f %7.0f %7.0f | %8.0f %7.0f | %7.0f%s" % (
        name, v[7], v[3] / tiles, v[0] / tiles, v[1] / tiles, v[2] / tiles, v[5] / tiles, v[4] / tiles, v[6] / tiles, extra))
