"""BASELINE configs[1]: LinkNet34 forward / forward+backward on DSB2018-shaped synthetic 256x256 batches (bf16, one B200).
CUDA events, warm; the optimiser step is not part of the reference's config line (torch_train.py:186-190 runs it, the
BASELINE config names forward/backward).  Prints ms per step and images/s; `--cpu` times the fp32 oracle on the host."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snb_b200  # noqa: E402,F401
from oracle import nets_oracle as no  # noqa: E402
from snb_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--cpu", action="store_true")
args = ap.parse_args()
sd = synth.linknet34_state_dict(seed=0)
rs = np.random.RandomState(0)
x = torch.from_numpy(rs.standard_normal((args.batch, 3, args.size, args.size)).astype(np.float32))
t = torch.from_numpy((rs.rand(args.batch, 1, args.size, args.size) > 0.5).astype(np.int64))
if args.cpu:
    torch.set_num_threads(os.cpu_count())
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    t0 = time.perf_counter()
    logits, _ = no.linknet34_forward_train(leaf, x)
    t1 = time.perf_counter()
    (no.bce_jaccard(logits, t) * args.batch).backward()
    t2 = time.perf_counter()
    print("cpu oracle (fp32, %d threads) batch %d x %d^2: forward %.1f ms, forward+backward %.1f ms = %.2f images/s" % (
        os.cpu_count(), args.batch, args.size, (t1 - t0) * 1e3, (t2 - t0) * 1e3, args.batch / (t2 - t0)))
    sys.exit(0)
from snb_b200.lib import losses  # noqa: E402
from snb_b200.lib.models import LinkNet34  # noqa: E402

m = LinkNet34(pretrained=False)
m.load_state_dict(sd)
m = m.cuda().train()
m.finaldrop1.p = 0.0
xd, td = x.cuda(), t.cuda()
crit = losses.BCEWithLogitsLossAndSmoothJaccard()


def fwd():
    with torch.no_grad():
        return m(xd)


def step():
    for p in m.parameters():
        p.grad = None
    loss = crit(m(xd), td) * args.batch
    loss.backward()
    return loss


def fused():
    return m.train_step(xd, td, crit)[0]


for name, fn in (("forward (train mode, batch statistics)", fwd), ("forward + loss + backward (autograd)", step),
                 ("forward + loss + backward (train_step)", fused)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print("LinkNet34 batch %d x %d^2 %-40s %8.2f ms/step  %8.1f images/s" % (args.batch, args.size, name, ms, args.batch / ms * 1e3))

plan = m.plan_train(args.batch, args.size, args.size)
flops = plan.flops + plan.bwd_flops
print("plan: %d forward launches, %d backward launches, %.1f GFLOP forward, %.1f GFLOP backward (algorithmic)" % (
    plan.launches, plan.bwd_launches, plan.flops / 1e9, plan.bwd_flops / 1e9))
