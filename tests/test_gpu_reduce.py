"""-m gpu: fused loss / IoU / PR-curve reductions through the C ABI against the oracle and reference values."""
import os

import numpy as np
import pytest
import torch

import snb_b200  # noqa: F401
from oracle import nets_oracle as no
from oracle import synth
from snb_b200.lib import losses, metrics

pytestmark = pytest.mark.gpu

REL = 5e-6   # float sums: fp32 elementwise with SFU exp/log/rcp, fp64 accumulation; reference sums in fp32 -> tolerance, not bits


@pytest.mark.parametrize("seed", [0, 3])
def test_against_reference_values(cuda, kats, seed):
    k = kats["loss"]["seed%d" % seed]
    logits, targets = synth.logits_targets(seed, tuple(k["shape"]))
    x, t = logits.cuda(), targets.cuda()
    assert float(losses.BCEWithLogitsLossAndSmoothJaccard()(x, t)) == pytest.approx(k["bce_jaccard"], rel=REL)
    assert float(losses.BCEWithSigmoidLoss()(x, t)) == pytest.approx(k["bce"], rel=REL)
    assert float(losses.SmoothJaccardLoss()(x, t)) == pytest.approx(k["smooth_jaccard"], rel=REL)
    assert float(metrics.JaccardScore()(x, t)) == pytest.approx(k["jaccard_score"], rel=REL)
    assert float(metrics.PixelAccuracy()(x, t)) == pytest.approx(k["pixel_accuracy"], rel=1e-7)
    assert metrics.confusion_counts(x, t).tolist() == k["counts"]                       # integer: bit-exact
    assert metrics.confusion_counts_from_probs(torch.sigmoid(logits).cuda(), t).tolist() == k["counts"]
    m = metrics.PRCurveMeter()
    m.update(x, t)
    assert m.tp.tolist() == k["pr_tp"] and m.tn.tolist() == k["pr_tn"]
    assert m.fp.tolist() == k["pr_fp"] and m.fn.tolist() == k["pr_fn"]
    m.update(x, t)                                                                      # accumulates like the reference
    assert m.tp.tolist() == [2 * v for v in k["pr_tp"]]


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1023, 4096 + 7])
@pytest.mark.parametrize("tdt", [torch.int64, torch.uint8, torch.float32])
def test_ragged_sizes_and_target_dtypes(cuda, n, tdt):
    logits, targets = synth.logits_targets(100 + n, (n,))
    s, c = losses.fused_sums(logits.cuda(), targets.to(tdt).cuda())
    p = torch.sigmoid(logits.double())
    z = torch.nn.functional.logsigmoid(logits.double())
    bce = torch.nn.functional.binary_cross_entropy_with_logits(z, targets.double(), reduction="sum")
    want = torch.stack([bce, (p * targets).sum(), p.sum(), targets.double().sum()])
    assert torch.allclose(s.cpu()[:4], want, rtol=1e-5, atol=1e-6) and float(s[4]) == 0.0     # no focal term requested
    assert c.tolist() == no.confusion_counts(torch.sigmoid(logits), targets).tolist()
    assert int(c.sum()) == n


def test_no_cpu_fallback():
    logits, targets = synth.logits_targets(0, (16,))
    with pytest.raises(RuntimeError):
        losses.fused_sums(logits, targets)


def test_full_size_counts_are_additive(cuda):
    """169 x 512 x 512 elements (one Inria image of tiles): counts of the halves add up to the whole (the multi-GPU
    all-reduce property) and equal torch's own integer count."""
    n = 169 * 512 * 512
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, device="cuda", generator=g)
    t = (torch.rand(n, device="cuda", generator=g) > 0.5).to(torch.uint8)
    whole = metrics.confusion_counts(x, t)
    h = n // 2
    parts = metrics.confusion_counts(x[:h], t[:h]) + metrics.confusion_counts(x[h:], t[h:])
    assert whole.tolist() == parts.tolist() and int(whole.sum()) == n
    pred = torch.sigmoid(x) > 0.5
    assert int(whole[0]) == int((pred & (t != 0)).sum()) and int(whole[3]) == int((~pred & (t == 0)).sum())


@pytest.mark.parametrize("seed,shape", [(0, (8, 1, 224, 224)), (3, (2, 1, 33, 17))])
@pytest.mark.parametrize("target_dtype", [torch.int64, torch.uint8, torch.float32])
def test_loss_backward_against_reference_gradients(cuda, golden_dir, seed, shape, target_dtype):
    """loss.backward() through the fused reduction + snb_loss_grad == torch autograd through the reference's modules
    (tests/golden/loss_grad.npz), with a non-unit upstream gradient, for all three losses."""
    from snb_b200.lib import losses

    g = np.load(os.path.join(golden_dir, "loss_grad.npz"))
    logits, targets = synth.logits_targets(seed, shape)
    t = targets.to(target_dtype).cuda()
    for name, mod in (("bce_jaccard", losses.BCEWithLogitsLossAndSmoothJaccard()),
                      ("smooth_jaccard", losses.SmoothJaccardLoss()), ("bce", losses.BCEWithSigmoidLoss())):
        x = logits.cuda().requires_grad_(True)
        loss = mod(x, t)
        assert loss.requires_grad and loss.dim() == 0
        (loss * 3.0).backward()
        got = x.grad.cpu().numpy().reshape(-1)
        want = g["seed%d_%s" % (seed, name)]
        got = got if got.size < 5000 else got[::97]
        assert x.grad.shape == logits.shape
        assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max(), name
    # without grad the forward-only path is taken and the value is the same
    with torch.no_grad():
        v = losses.BCEWithLogitsLossAndSmoothJaccard()(logits.cuda(), t)
    x = logits.cuda().requires_grad_(True)
    assert float(v) == pytest.approx(float(losses.BCEWithLogitsLossAndSmoothJaccard()(x, t).detach()), rel=1e-6)


@pytest.mark.parametrize("seed,shape", [(0, (8, 1, 224, 224)), (3, (2, 1, 33, 17))])
@pytest.mark.parametrize("target_dtype", [torch.int64, torch.uint8, torch.float32])
def test_extra_losses_against_reference(cuda, golden_dir, seed, shape, target_dtype):
    """JaccardLoss, FocalLossBinary (gamma 2 / 1.5 / 0, mean / sum) and BCEWithSigmoidLoss (sum, reduce=False) on the
    fused reduction: values and gradients against the reference modules (tests/golden/loss_extra.npz)."""
    g = np.load(os.path.join(golden_dir, "loss_extra.npz"))
    logits, targets = synth.logits_targets(seed, shape)
    t = targets.to(target_dtype).cuda()
    sample = lambda a: a if a.size < 5000 else a[::97]
    up = torch.from_numpy(np.random.RandomState(40 + seed).standard_normal(shape).astype(np.float32)).cuda()
    cases = [("jaccard", losses.JaccardLoss(), None, REL),
             ("focal_g2_mean", losses.FocalLossBinary(gamma=2), None, 2e-5),
             ("focal_g1.5_sum", losses.FocalLossBinary(gamma=1.5, size_average=False), None, 2e-5),
             ("focal_g0_mean", losses.FocalLossBinary(gamma=0), None, REL),
             ("bce_sum", losses.BCEWithSigmoidLoss(size_average=False), None, REL),
             ("bce_elem", losses.BCEWithSigmoidLoss(reduce=False), up, None)]
    for tag, mod, upstream, rel in cases:
        with torch.no_grad():
            v = mod(logits.cuda(), t)
        x = logits.cuda().requires_grad_(True)
        y = mod(x, t)
        assert y.requires_grad
        if upstream is None:
            assert y.dim() == 0 and float(y) == pytest.approx(float(g["seed%d_%s" % (seed, tag)]), rel=rel), tag
            assert float(v) == pytest.approx(float(y.detach()), rel=1e-6)
            (y * 3.0).backward()
        else:
            assert y.shape == logits.shape and torch.equal(v, y.detach())
            want_e = g["seed%d_%s" % (seed, tag)]
            assert np.abs(sample(y.detach().cpu().numpy().reshape(-1)) - want_e).max() < 2e-6 * max(1.0, np.abs(want_e).max())
            (y * upstream).sum().backward()
        want = g["seed%d_%s_grad" % (seed, tag)]
        got = sample(x.grad.cpu().numpy().reshape(-1))
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max() + 1e-12, (tag, np.abs(got - want).max(), np.abs(want).max())


def test_reductions_are_single_launch_and_deterministic(cuda):
    """The workspace returns to rest after every call (back-to-back calls on one workspace agree bit for bit), and
    two streams use two workspaces."""
    logits, targets = synth.logits_targets(5, (3, 1, 300, 301))
    x, t = logits.cuda(), targets.cuda()
    a = [losses.fused_sums(x, t, focal_gamma=2.0) for _ in range(3)]
    for s, c in a[1:]:
        assert torch.equal(s, a[0][0]) and torch.equal(c, a[0][1])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        s2, c2 = losses.fused_sums(x, t, focal_gamma=2.0)
    side.synchronize()
    assert torch.equal(s2, a[0][0]) and torch.equal(c2, a[0][1])
    from snb_b200 import _native as N
    ws = N.reduce_workspace()
    torch.cuda.synchronize()
    assert int(ws[:64].count_nonzero()) == 0        # the ticket is back at rest
