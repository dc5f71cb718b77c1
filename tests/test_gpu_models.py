"""-m gpu: UNet16 / UNet11 forward on the native engine against the reference's logits and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

import snb_b200  # noqa: F401
from oracle import nets_oracle as no
from oracle import synth
from snb_b200.lib.models import UNet11, UNet16

pytestmark = pytest.mark.gpu

BF16_PROB_TOL = 2e-2   # north-star: probabilities within 2e-2 abs in bf16 mode


@pytest.mark.parametrize("arch,cls", [("unet16", UNet16), ("unet11", UNet11)])
def test_logits_against_reference_vectors(cuda, golden_dir, arch, cls):
    g = np.load(os.path.join(golden_dir, "models.npz"))
    m = cls()
    m.load_state_dict(synth.vgg_unet_state_dict(arch, seed=1), strict=True)   # the reference's key names
    m = m.cuda().eval()
    x = torch.from_numpy(g[arch + "_x"]).cuda()
    with torch.no_grad():
        y = m(x)
    assert y.shape == (2, 1, 64, 96) and y.dtype == torch.float32
    ref = torch.from_numpy(g[arch + "_logits"])
    p_err = (torch.sigmoid(y.cpu()) - torch.sigmoid(ref)).abs().max().item()
    assert p_err < BF16_PROB_TOL, p_err
    # against the oracle evaluated with bf16-rounded conv operands the agreement is much tighter
    sd = synth.vgg_unet_state_dict(arch, seed=1)
    with torch.no_grad():
        q = no.unet_vgg_forward(sd, torch.from_numpy(g[arch + "_x"]), arch, quant=no.bf16_round)
    l_err = (y.cpu() - q).abs().max().item()
    assert l_err < 0.02 * max(1.0, q.abs().max().item()), l_err


def test_no_cpu_fallback():
    m = UNet16()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 32, 32))


def test_plan_tracks_weight_updates(cuda):
    m = UNet16().cuda().eval()
    x = torch.randn(1, 3, 32, 32, device="cuda")
    with torch.no_grad():
        a = m(x)
        m.final.bias.add_(1.0)
        b = m(x)
    assert torch.allclose(b - a, torch.ones_like(a), atol=1e-5)
