"""Per-parameter gradient errors of the LinkNet34 training step against the fp32 CPU oracle (debugging aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snb_b200  # noqa: E402,F401
from oracle import nets_oracle as no  # noqa: E402
from snb_b200 import synth  # noqa: E402
from snb_b200.lib import losses  # noqa: E402
from snb_b200.lib.models import LinkNet34  # noqa: E402

sd = synth.linknet34_state_dict(seed=6)
n, hw = 8, int(sys.argv[1]) if len(sys.argv) > 1 else 128
rs = np.random.RandomState(31)
x = torch.from_numpy(rs.standard_normal((n, 3, hw, hw)).astype(np.float32))
t = torch.from_numpy((rs.rand(n, 1, hw, hw) > 0.5).astype(np.int64))
leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
quant = no.bf16_round if os.environ.get("ORACLE_BF16", "1") == "1" else None     # same rounding points as the device path
linear = os.environ.get("LINEAR", "0") == "1"
logits_ref, _ = no.linknet34_forward_train(leaf, x, quant=quant, linear=linear)
(no.bce_jaccard(logits_ref, t) * n).backward()
m = LinkNet34(pretrained=False)
m.load_state_dict(sd)
m = m.cuda().train()
m.finaldrop1.p = 0.0
m._test_linear = linear
logits = m(x.cuda())
(losses.BCEWithLogitsLossAndSmoothJaccard()(logits, t.cuda()) * n).backward()
print("logit err", (logits.detach().cpu() - logits_ref.detach()).abs().max().item())
for name, p in m.named_parameters():
    want, got = leaf[name].grad, p.grad.cpu()
    print("%-40s ref %.3e got %.3e rel-L2 %.4f" % (name, want.norm().item(), got.norm().item(),
                                                   ((got - want).norm() / (want.norm() + 1e-20)).item()))
