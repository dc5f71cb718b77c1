"""-m gpu: InPlaceABN (lib/modules/abn) on the native kernels against the reference's vectors and the oracle."""
import os

import numpy as np
import pytest
import torch

import snb_b200  # noqa: F401
from oracle import nets_oracle as no
from snb_b200.lib.modules.abn import InPlaceABN, inplace_abn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_against_reference_vectors(cuda, golden_dir, mode):
    g = np.load(os.path.join(golden_dir, "abn.npz"))
    training = mode == "train"
    m = InPlaceABN(6).cuda()
    with torch.no_grad():
        m.weight.copy_(torch.from_numpy(g["weight"]))
        m.bias.copy_(torch.from_numpy(g["bias"]))
        m.running_mean.copy_(torch.from_numpy(g["running_mean"]))
        m.running_var.copy_(torch.from_numpy(g["running_var"]))
    m.train(training)
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    z = m(x.clone())                                   # in place on the clone, like conv -> abn in DecoderBlockLinkNet
    zc = z.detach().clone()
    z.backward(torch.from_numpy(g["grad"]).cuda())
    assert np.abs(zc.cpu().numpy() - g[mode + "_z"]).max() < 1e-5
    assert np.abs(m.running_mean.cpu().numpy() - g[mode + "_running_mean"]).max() < 1e-6
    assert np.abs(m.running_var.cpu().numpy() - g[mode + "_running_var"]).max() < 1e-6
    assert np.abs(x.grad.cpu().numpy() - g[mode + "_dx"]).max() < 1e-5
    assert np.abs(m.weight.grad.cpu().numpy() - g[mode + "_dweight"]).max() < 2e-4
    assert np.abs(m.bias.grad.cpu().numpy() - g[mode + "_dbias"]).max() < 2e-4


@pytest.mark.parametrize("shape,activation,affine", [
    ((8, 64, 64, 64), "leaky_relu", True),       # LinkNet34 decoder shapes (float4 path)
    ((4, 128, 32, 32), "leaky_relu", True),
    ((2, 16, 17, 23), "leaky_relu", True),       # H*W not a multiple of 4: scalar path
    ((3, 10, 9, 9), "elu", True),
    ((2, 8, 12, 12), "none", False),
    ((16, 32), "leaky_relu", True),              # 2-D input (count = N)
])
@pytest.mark.parametrize("training", [True, False])
def test_against_oracle(cuda, shape, activation, affine, training):
    gen = torch.Generator().manual_seed(len(shape) * 100 + shape[1])
    c = shape[1]
    x = torch.randn(shape, generator=gen) * 2.0 + 0.5
    dz = torch.randn(shape, generator=gen)
    w = (torch.rand(c, generator=gen) + 0.5) * torch.where(torch.rand(c, generator=gen) < 0.3, -1.0, 1.0) if affine else None
    b = torch.randn(c, generator=gen) * 0.3 if affine else None
    rm, rv = torch.randn(c, generator=gen) * 0.1, torch.rand(c, generator=gen) + 0.5
    x4 = x if x.dim() == 4 else x.view(shape[0], c, 1, 1)
    z_ref, var_ref, rm_ref, rv_ref = no.inplace_abn_forward(x4, w, b, rm, rv, training, 0.1, 1e-5, activation, 0.01)
    dx_ref, dw_ref, db_ref = no.inplace_abn_backward(z_ref, dz.view(x4.shape), var_ref, w, b, training, 1e-5, activation, 0.01)
    xd = x.cuda().requires_grad_(True)
    wd = w.cuda().requires_grad_(True) if affine else None
    bd = b.cuda().requires_grad_(True) if affine else None
    rmd, rvd = rm.cuda(), rv.cuda()
    z = inplace_abn(xd.clone(), wd, bd, rmd, rvd, training, 0.1, 1e-5, activation, 0.01)
    zc = z.detach().clone()
    z.backward(dz.cuda())
    scale = max(1.0, z_ref.abs().max().item())
    assert (zc.cpu().view(x4.shape) - z_ref).abs().max().item() < 2e-5 * scale
    assert (rmd.cpu() - rm_ref).abs().max().item() < 1e-5 and (rvd.cpu() - rv_ref).abs().max().item() < 1e-4
    assert (xd.grad.cpu().view(x4.shape) - dx_ref).abs().max().item() < 1e-4 * max(1.0, dx_ref.abs().max().item())
    if affine:   # float32 sums of up to 32k terms in the oracle against float64 accumulation on the device
        assert (wd.grad.cpu() - dw_ref).abs().max().item() < 3e-4 * max(1.0, dw_ref.abs().max().item())
        assert (bd.grad.cpu() - db_ref).abs().max().item() < 3e-4 * max(1.0, db_ref.abs().max().item())


def test_argument_errors(cuda):
    m = InPlaceABN(4).cuda()
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 4, 3, 3))                                   # CPU tensor: no fallback
    with pytest.raises(ValueError):
        m(torch.zeros(2, 4, 3, 3, device="cuda", dtype=torch.float16))
    with pytest.raises(ValueError):
        inplace_abn(torch.zeros(2, 4, 3, 3, device="cuda"), m.weight, m.bias, m.running_mean, m.running_var, True, 0.1,
                    1e-5, "swish", 0.01)


def test_large_mean_and_non_contiguous_input(cuda):
    """Batch variance of a channel whose |mean| is 1000x its standard deviation (E[x^2] - E[x]^2 in float32 would lose
    every bit; the kernels accumulate around a pivot), on a non-contiguous input (the reference copies it,
    lib/modules/abn/functions.py:72), for the NCHW module and the NHWC slab kernel."""
    from snb_b200 import engine as E
    from snb_b200 import _native as N

    gen = torch.Generator().manual_seed(11)
    n, c, h, w = 4, 16, 24, 24
    x = torch.randn((n, h, w, c), generator=gen) * 0.05 + 50.0
    xd = x.cuda().permute(0, 3, 1, 2)                       # NCHW view of NHWC storage: not contiguous
    assert not xd.is_contiguous()
    m = InPlaceABN(c).cuda().train()
    z = m(xd)
    ref = x.permute(0, 3, 1, 2).double()
    mean, var = ref.mean(dim=(0, 2, 3)), ref.var(dim=(0, 2, 3), unbiased=False)
    want = torch.nn.functional.leaky_relu((ref - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + 1e-5) * (1 + 1e-5), 0.01)
    assert (z.cpu().double() - want).abs().max().item() < 2e-3          # float32 input resolution at 50 is 4e-6 = 1e-4 sigma
    cnt = n * h * w
    assert (m.running_var.cpu().double() - (0.9 + 0.1 * var * cnt / (cnt - 1))).abs().max().item() < 1e-6
    # NHWC bf16 slab kernel: values 50 +- 0.25 in bf16 (step 0.25): variance of the rounded data must be recovered
    src = E.Slab(n, h, w, c, "cuda")
    src.t.copy_((torch.randn((n, h, w, c), generator=gen) * 0.5 + 50.0).cuda().to(torch.bfloat16))
    dst = E.Slab(n, h, w, c, "cuda")
    ones, zeros = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    op = E.BnTrainOp(src.view(), dst.view(), (ones, zeros, zeros.clone(), ones.clone(), 1e-5, 0.1), False, -1.0)
    op(N.stream_ptr())
    xs = src.t.double().cpu()
    assert (op.mean.cpu().double() - xs.mean(dim=(0, 1, 2))).abs().max().item() < 1e-4
    v = xs.var(dim=(0, 1, 2), unbiased=False)
    assert ((op.var.cpu().double() - v).abs() / v).max().item() < 1e-4
