"""CPU dry run of the training plan's HOST logic (no GPU needed): the native library is replaced by a stub whose entry
points all succeed, slabs live in host memory, and the plan is built and its op lists walked once.  Catches Python-side
mistakes (argument lists, bookkeeping of gradient slabs, index maps) before spending GPU time."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snb_b200  # noqa: E402,F401
from snb_b200 import _native as N  # noqa: E402


class _Stub:
    def __getattr__(self, name):
        if name.endswith("_flops"):
            return lambda *a: 1.0
        if name == "snb_reduce_workspace_bytes":
            return lambda *a: 1 << 18
        return lambda *a: 0


N._lib = _Stub()
N.require_cuda = lambda: None
N.stream_ptr = lambda: N.c_vp(0)
N.ptr = lambda t: N.c_vp(0 if t is None else t.data_ptr())

from snb_b200.lib.models import LinkNet34  # noqa: E402
from snb_b200.train_engine import LinkNet34TrainPlan  # noqa: E402

m = LinkNet34(pretrained=False).train()
plan = LinkNet34TrainPlan(m, 2, 64, 96, torch.device("cpu"))
plan.use_graph = False
plan.load_nchw(torch.zeros(2, 3, 64, 96))
plan.run()
g = plan.backward(torch.zeros(2, 1, 64, 96))
assert set(g) == set(m.parameters())
print("forward ops %d (%d launches), backward ops %d (%d launches); repack segments %d + %d, unpack segments %d; "
      "gradient arena %.1f MB" % (len(plan.ops), plan.launches, len(plan.bwd_ops), plan.bwd_launches, len(plan.repack.rows),
                                  len(plan.repack32.rows), len(plan.unpack.rows), plan.grad_arena.numel() * 4 / 1e6))
kinds = {}
for op in plan.bwd_ops:
    kinds[type(op).__name__] = kinds.get(type(op).__name__, 0) + 1
print(kinds)
