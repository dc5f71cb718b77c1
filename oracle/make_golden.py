"""Generate tests/golden/* by RUNNING THE REFERENCE (imported from /root/reference) on seeded synthetic inputs.

Runs only in the build container (the reference is not present on the GPU box).  The reference ships no golden
vectors of its own, so these files are what pins the oracle and, through it, the CUDA path.
Usage: python oracle/make_golden.py
"""
import collections
import collections.abc
import json
import os
import sys
import types
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SNB_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
from oracle import synth  # noqa: E402

# non-invasive import shims: optional deps the hot path never uses
for name in ("tensorboardX", "matplotlib", "matplotlib.pyplot", "seaborn"):
    sys.modules.setdefault(name, types.ModuleType(name))
collections.Iterable = collections.abc.Iterable

def _install_inplace_abn_shim():
    """Pure-torch stand-in for the un-vendored `inplace_abn` CUDA extension (mapillary/inplace_abn, no pinned version),
    implementing the calls lib/modules/abn/functions.py makes on the eval path with that library's published forward:
    y = (x - mean) / sqrt(var + eps) * (|weight| + eps) + bias, in place; leaky_relu in place."""
    m = types.ModuleType("inplace_abn")

    def forward(x, mean, var, weight, bias, affine, eps):
        g = (weight.abs() + eps) if affine else torch.ones_like(mean)
        b = bias if affine else torch.zeros_like(mean)
        x.sub_(mean.view(1, -1, 1, 1)).div_(torch.sqrt(var.view(1, -1, 1, 1) + eps)).mul_(g.view(1, -1, 1, 1)).add_(b.view(1, -1, 1, 1))
        return True

    def leaky_relu_forward(x, slope):
        x.copy_(torch.nn.functional.leaky_relu(x, slope))
        return True

    def mean_var(x):
        return x.mean(dim=(0, 2, 3)), x.var(dim=(0, 2, 3), unbiased=False)

    # backward half of the backend as published (inplace_abn_cuda.cu: leaky_relu_backward, edz_eydz, backward)
    def leaky_relu_backward(z, dz, slope):
        neg = z < 0
        dz[neg] *= slope
        z[neg] /= slope
        return True

    def _gamma_beta(z, weight, bias, affine, eps):
        c = z.shape[1]
        g = (weight.abs() + eps) if affine else torch.ones(c)
        b = bias if affine else torch.zeros(c)
        return g.view(1, -1, 1, 1), b.view(1, -1, 1, 1)

    def edz_eydz(z, dz, weight, bias, affine, eps):
        g, b = _gamma_beta(z, weight, bias, affine, eps)
        y = (z - b) / g
        return dz.sum(dim=(0, 2, 3)), (y * dz).sum(dim=(0, 2, 3))

    def backward(z, dz, var, weight, bias, edz, eydz, affine, eps):
        g, b = _gamma_beta(z, weight, bias, affine, eps)
        count = z.numel() // z.shape[1]
        y = (z - b) / g
        mul = g * torch.rsqrt(var.view(1, -1, 1, 1) + eps)
        dx = (dz - (edz / count).view(1, -1, 1, 1) - y * (eydz / count).view(1, -1, 1, 1)) * mul
        dweight = torch.where(weight > 0, eydz, -eydz) if affine else dz.new_empty(0)
        dbias = edz if affine else dz.new_empty(0)
        return dx, dweight, dbias

    m.forward, m.leaky_relu_forward, m.mean_var = forward, leaky_relu_forward, mean_var
    m.leaky_relu_backward, m.edz_eydz, m.backward = leaky_relu_backward, edz_eydz, backward
    sys.modules["inplace_abn"] = m


_install_inplace_abn_shim()

from lib import augmentations as aug  # noqa: E402
from lib import losses as ref_losses  # noqa: E402
from lib import metrics as ref_metrics  # noqa: E402
from lib.common import InMemoryDataset  # noqa: E402
from lib.datasets.Inria import INRIA_MEAN, INRIA_STD  # noqa: E402
from lib.models.unet11 import UNet11  # noqa: E402
from lib.models.unet16 import UNet16  # noqa: E402
from lib.models.zf_unet import ZF_UNET  # noqa: E402
from lib.models.tiramisu import FCDenseNet67  # noqa: E402
from lib.models.linknet import LinkNet34  # noqa: E402
from lib.tiles import ImageSlicer, compute_patch_weight_loss  # noqa: E402
from lib.train_utils import PRCurveMeter  # noqa: E402


def slicer_kats():
    cases = [((5000, 5000, 3), 512, 384, 0), ((5000, 5000, 3), 512, 256, 0), ((5000, 5000, 3), 224, 112, 0),
             ((5000, 5000), 1024, 512, 0), ((360, 360), 256, 128, 0), ((256, 256), 256, 128, 0),
             ((50, 70), 256, 128, 0), ((5000, 5000), 512, 384, 60), ((37, 53, 3), 16, 8, 0), ((37, 53), 16, 16, 0),
             ((100, 90), 32, 20, 0), ((96, 80, 3), 64, 32, 0), ((5, 7), 32, 16, 0), ((1, 9), 4, 2, 0)]
    out = []
    for shape, tile, step, margin in cases:
        s = ImageSlicer(shape, tile, step, image_margin=margin)
        out.append(dict(shape=list(shape), tile=tile, step=step, margin=margin,
                        margins=[s.margin_left, s.margin_right, s.margin_top, s.margin_bottom],
                        n_crops=len(s.crops), crops_head=[list(c) for c in s.crops[:3]],
                        crops_tail=[list(c) for c in s.crops[-3:]],
                        crops_sum=[int(sum(c[0] for c in s.crops)), int(sum(c[1] for c in s.crops))]))
    errors = []
    for shape, tile, step, margin in [((5000, 5000), 512, 384, 10), ((5000, 5000), 512, 0, 0),
                                      ((64, 64), 32, 33, 0), ((64, 64), 32, -1, 0)]:
        try:
            ImageSlicer(shape, tile, step, image_margin=margin)
            errors.append(dict(shape=list(shape), tile=tile, step=step, margin=margin, error=None))
        except Exception as e:  # noqa: BLE001
            errors.append(dict(shape=list(shape), tile=tile, step=step, margin=margin, error=type(e).__name__))
    return dict(cases=out, errors=errors)


def split_merge_vectors():
    rs = np.random.RandomState(7)
    d = {}
    # (name, shape, dtype, tile, step)
    specs = [("u8c3", (37, 53, 3), np.uint8, 16, 8), ("f32c1", (40, 33, 1), np.float32, 16, 12),
             ("u8_2d", (37, 53), np.uint8, 16, 16), ("f64c3", (21, 30, 3), np.float64, 8, 4),
             ("tiny_multi_reflect", (5, 7, 1), np.float32, 32, 16)]
    for name, shape, dt, tile, step in specs:
        img = (rs.randint(0, 256, shape).astype(dt) if dt == np.uint8 else rs.standard_normal(shape).astype(dt))
        s = ImageSlicer(shape, tile, step)
        tiles = s.split(img)
        d[name + "_image"] = img
        d[name + "_tiles"] = np.stack(tiles)
        d[name + "_cfg"] = np.array([tile, step])
        d[name + "_cut3"] = s.cut_patch(img, min(3, len(s.crops) - 1))
    for weight in ("mean", "pyramid"):
        s = ImageSlicer((37, 53, 3), 16, 8, weight=weight)
        tiles = [rs.rand(16, 16, 2).astype(np.float32) for _ in s.crops]
        d["merge_%s_tiles" % weight] = np.stack(tiles)
        d["merge_%s_out" % weight] = s.merge(tiles)
        # identity: merge(split(x)) == x exactly
        img = rs.rand(37, 53, 3).astype(np.float32)
        d["merge_%s_identity_in" % weight] = img
        d["merge_%s_identity_out" % weight] = s.merge(s.split(img))
    s = ImageSlicer((37, 53, 3), 16, 8, weight="mean")
    tiles = [rs.randint(0, 256, (16, 16, 3)).astype(np.uint8) for _ in s.crops]
    d["merge_u8_tiles"] = np.stack(tiles)
    d["merge_u8_out"] = s.merge(tiles, dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "split_merge.npz"), **d)


def weight_vectors():
    w16 = compute_patch_weight_loss(16, 16)[0]
    w24 = compute_patch_weight_loss(24, 24)[0]
    w64 = compute_patch_weight_loss(64, 64)[0]
    np.savez_compressed(os.path.join(OUT, "pyramid.npz"), w16=w16, w24=w24, w64=w64)
    # survey KAT for the 512 tile (7 s python loop in the reference): min / max / sum, SURVEY 8c'
    return dict(n512=dict(min=0.006262400053660891, max=3.1974996323919567, sum=262143.99999999997),
                n64=dict(min=float(w64.min()), max=float(w64.max()), sum=float(w64.sum())))


def normalize_vectors():
    t = aug.NormalizeImage(mean=INRIA_MEAN, std=INRIA_STD)
    levels = np.repeat(np.arange(256, dtype=np.uint8)[:, None, None], 3, axis=2)   # 256 x 1 x 3 BGR "image"
    out64 = t(levels)
    ds = InMemoryDataset([out64], None)
    np.savez_compressed(os.path.join(OUT, "normalize.npz"), levels=levels, out64=out64, chw_f32=ds[0].numpy())


def tta_vectors():
    rs = np.random.RandomState(11)
    tiles = [rs.rand(6, 6, 2).astype(np.float32) for _ in range(2)]
    views = aug.tta_d4_aug(tiles)
    preds = [rs.rand(6, 6, 1).astype(np.float32) for _ in range(16)]
    deaug = aug.tta_d4_deaug(preds)
    np.savez_compressed(os.path.join(OUT, "tta.npz"), tiles=np.stack(tiles), views=np.stack(views),
                        preds=np.stack(preds), deaug=np.stack(deaug))


def loss_vectors():
    out = {}
    for seed, shape in [(0, (8, 1, 224, 224)), (3, (2, 1, 33, 17))]:
        logits, targets = synth.logits_targets(seed, shape)
        loss = ref_losses.BCEWithLogitsLossAndSmoothJaccard()
        loss.bce_loss.size_average = True   # attributes modern _Loss no longer stores
        loss.bce_loss.reduce = True
        bce = ref_losses.BCEWithSigmoidLoss.__new__(ref_losses.BCEWithSigmoidLoss)
        torch.nn.Module.__init__(bce)
        bce.size_average, bce.reduce = True, True
        pa = ref_metrics.PixelAccuracy()(logits, targets)
        meter = PRCurveMeter()
        meter.update(logits, targets)
        p = torch.sigmoid(logits)
        pred = p > 0.5
        t = targets.bool()
        out["seed%d" % seed] = dict(
            shape=list(shape),
            bce_jaccard=float(loss(logits, targets)), bce=float(bce(logits, targets)),
            smooth_jaccard=float(ref_losses.SmoothJaccardLoss()(logits, targets)),
            jaccard_score=float(ref_metrics.JaccardScore()(logits, targets)), pixel_accuracy=float(pa),
            counts=[int((pred & t).sum()), int((pred & ~t).sum()), int((~pred & t).sum()), int((~pred & ~t).sum())],
            pr_tp=[int(v) for v in meter.tp], pr_tn=[int(v) for v in meter.tn],
            pr_fp=[int(v) for v in meter.fp], pr_fn=[int(v) for v in meter.fn])
    # gradients of the three losses with respect to the logits (torch autograd through the reference's modules)
    grads = {}
    for seed, shape in [(0, (8, 1, 224, 224)), (3, (2, 1, 33, 17))]:
        logits, targets = synth.logits_targets(seed, shape)
        for name, mod in (("bce_jaccard", ref_losses.BCEWithLogitsLossAndSmoothJaccard()),
                          ("smooth_jaccard", ref_losses.SmoothJaccardLoss()),
                          ("bce", ref_losses.BCEWithSigmoidLoss.__new__(ref_losses.BCEWithSigmoidLoss))):
            if name == "bce":
                torch.nn.Module.__init__(mod)
                mod.size_average, mod.reduce = True, True
            if name == "bce_jaccard":
                mod.bce_loss.size_average, mod.bce_loss.reduce = True, True
            x = logits.clone().requires_grad_(True)
            (mod(x, targets) * 3.0).backward()          # a non-unit upstream gradient
            g = x.grad.numpy().reshape(-1)
            grads["seed%d_%s" % (seed, name)] = g if g.size < 5000 else g[::97].copy()
    np.savez_compressed(os.path.join(OUT, "loss_grad.npz"), **grads)
    return out


def loss_extra_vectors():
    """JaccardLoss, FocalLossBinary and BCEWithSigmoidLoss(reduce=False) of the reference (lib/losses.py:18-28,46-53,
    78-101): values and gradients with respect to the logits -> tests/golden/loss_extra.npz."""
    out = {}
    for seed, shape in [(0, (8, 1, 224, 224)), (3, (2, 1, 33, 17))]:
        logits, targets = synth.logits_targets(seed, shape)
        sample = (lambda a: a if a.size < 5000 else a[::97].copy())

        def run(tag, mod, upstream=None):
            x = logits.clone().requires_grad_(True)
            y = mod(x, targets)
            if upstream is None:
                out["seed%d_%s" % (seed, tag)] = np.float64(float(y))
                (y * 3.0).backward()
            else:
                out["seed%d_%s" % (seed, tag)] = sample(y.detach().numpy().reshape(-1))
                (y * upstream).sum().backward()
            out["seed%d_%s_grad" % (seed, tag)] = sample(x.grad.numpy().reshape(-1))

        run("jaccard", ref_losses.JaccardLoss())
        for tag, gamma, avg in (("focal_g2_mean", 2, True), ("focal_g1.5_sum", 1.5, False), ("focal_g0_mean", 0, True)):
            m = ref_losses.FocalLossBinary.__new__(ref_losses.FocalLossBinary)
            torch.nn.Module.__init__(m)
            m.gamma, m.size_average, m.reduce = gamma, avg, True
            run(tag, m)
        bce = ref_losses.BCEWithSigmoidLoss.__new__(ref_losses.BCEWithSigmoidLoss)
        torch.nn.Module.__init__(bce)
        bce.size_average, bce.reduce = True, False
        up = torch.from_numpy(np.random.RandomState(40 + seed).standard_normal(shape).astype(np.float32))
        # (the upstream gradient is regenerated from RandomState(40 + seed) by the tests)
        run("bce_elem", bce, upstream=up)
        bce.size_average = False
        bce.reduce = True
        run("bce_sum", bce)
    np.savez_compressed(os.path.join(OUT, "loss_extra.npz"), **out)


def model_vectors():
    d = {}
    for arch, cls in (("unet16", UNet16), ("unet11", UNet11)):
        m = cls()
        sd = synth.vgg_unet_state_dict(arch, seed=1)
        missing = m.load_state_dict(sd, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        m.eval()
        x = torch.from_numpy(np.random.RandomState(5).standard_normal((2, 3, 64, 96)).astype(np.float32))
        with torch.no_grad():
            d[arch + "_logits"] = m(x).numpy()
        d[arch + "_x"] = x.numpy()
    np.savez_compressed(os.path.join(OUT, "models.npz"), **d)


def zf_unet_vectors():
    """BASELINE configs[0]: ZF_UNET 224 forward + BCE-soft-Jaccard loss + IoU on a synthetic 8x3x224x224 batch (eval),
    plus a small 2x3x64x96 case whose full logits are stored."""
    m = ZF_UNET()
    sd = synth.zf_unet_state_dict(seed=4)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    m.eval()
    d = {}
    x_small = torch.from_numpy(np.random.RandomState(6).standard_normal((2, 3, 64, 96)).astype(np.float32))
    x, y = synth.logits_targets(8, (8, 1, 224, 224))
    x224 = torch.from_numpy(np.random.RandomState(8).standard_normal((8, 3, 224, 224)).astype(np.float32))
    with torch.no_grad():
        d["small_x"] = x_small.numpy()
        d["small_logits"] = m(x_small).numpy()
        logits = m(x224)
    d["cfg1_logits_sample"] = logits[:, :, ::7, ::7].numpy()          # 8 x 1 x 32 x 32 subsample of the 224 logits
    loss = ref_losses.BCEWithLogitsLossAndSmoothJaccard()
    loss.bce_loss.size_average = True
    loss.bce_loss.reduce = True
    pred = torch.sigmoid(logits) > 0.5
    t = y.bool()
    np.savez_compressed(os.path.join(OUT, "zf_unet.npz"), **d)
    return dict(bce_jaccard=float(loss(logits, y)), jaccard_score=float(ref_metrics.JaccardScore()(logits, y)),
                pixel_accuracy=float(ref_metrics.PixelAccuracy()(logits, y)),
                counts=[int((pred & t).sum()), int((pred & ~t).sum()), int((~pred & t).sum()), int((~pred & ~t).sum())],
                logit_min=float(logits.min()), logit_max=float(logits.max()))


def fcdensenet_vectors():
    """BASELINE configs[4] model: FCDenseNet67(n_classes=1) in eval mode with randomised BatchNorm buffers."""
    m = FCDenseNet67(n_classes=1)
    sd = synth.fcdensenet_state_dict(seed=5)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys and len(m.state_dict()) == 434
    m.eval()
    x = torch.from_numpy(np.random.RandomState(12).standard_normal((2, 3, 64, 96)).astype(np.float32))
    x224 = torch.from_numpy(np.random.RandomState(13).standard_normal((1, 3, 224, 224)).astype(np.float32))
    with torch.no_grad():
        y, y224 = m(x).numpy(), m(x224).numpy()
    np.savez_compressed(os.path.join(OUT, "fcdensenet67.npz"), x=x.numpy(), logits=y, logits224=y224)
    return dict(params=int(sum(p.numel() for p in m.parameters())), logit_min=float(y.min()), logit_max=float(y.max()))


def linknet34_vectors():
    """LinkNet34 (eval mode; InPlaceABN through the pure-torch shim of the un-vendored backend: parity unpinned)."""
    m = LinkNet34(pretrained=False)
    sd = synth.linknet34_state_dict(seed=6)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    m.eval()
    x = torch.from_numpy(np.random.RandomState(14).standard_normal((2, 3, 64, 96)).astype(np.float32))
    x256 = torch.from_numpy(np.random.RandomState(15).standard_normal((1, 3, 256, 256)).astype(np.float32))
    with torch.no_grad():
        y, y256 = m(x).numpy(), m(x256).numpy()
    # train() mode with Dropout2d off (BASELINE configs[1] as specified in SURVEY 8d): batch statistics everywhere.  torch >= 2
    # rejects the reference autograd function's mark_dirty(x, running_mean, running_var), so `inplace_abn` inside
    # lib.modules.abn.bn is pointed at the reference's own InPlaceABN.forward body with a stand-in context (no edit of the
    # reference; forward only)
    import lib.modules.abn.bn as ref_bn
    from lib.modules.abn.functions import InPlaceABN as RefABN

    class Ctx:
        def mark_dirty(self, *t):
            pass

        def save_for_backward(self, *t):
            self.saved_tensors = t

    orig = ref_bn.inplace_abn
    ref_bn.inplace_abn = lambda *a: RefABN.forward(Ctx(), *a)
    m.train()
    m.finaldrop1.p = 0.0
    xt = torch.from_numpy(np.random.RandomState(16).standard_normal((4, 3, 64, 64)).astype(np.float32))
    with torch.no_grad():
        yt = m(xt).numpy()
    ref_bn.inplace_abn = orig
    sd_after = m.state_dict()
    stats = {k.replace('.', '__'): sd_after[k].numpy() for k in (
        'firstbn.running_mean', 'firstbn.running_var', 'encoder2.0.downsample.1.running_var', 'encoder4.2.bn2.running_mean',
        'decoder4.abn1.running_var', 'decoder1.abn2.running_mean', 'decoder1.abn3.running_var')}
    np.savez_compressed(os.path.join(OUT, "linknet34.npz"), x=x.numpy(), logits=y, logits256=y256, train_x=xt.numpy(),
                        train_logits=yt, **stats)
    return dict(params=int(sum(p.numel() for p in m.parameters())), n_keys=len(m.state_dict()),
                logit_min=float(y.min()), logit_max=float(y.max()))


def linknet34_dropout_vectors():
    """The reference LinkNet34 in train() mode with its default Dropout2d(p=0.5) ACTIVE (lib/models/linknet.py:57,83): the
    keep mask torch drew is captured with a forward hook on `finaldrop1`, so a restatement given the same mask must
    reproduce the logits -> tests/golden/linknet34_dropout.npz."""
    import lib.modules.abn.bn as ref_bn
    from lib.modules.abn.functions import InPlaceABN as RefABN

    class Ctx:
        def mark_dirty(self, *t):
            pass

        def save_for_backward(self, *t):
            self.saved_tensors = t

    m = LinkNet34(pretrained=False)
    m.load_state_dict(synth.linknet34_state_dict(seed=6), strict=True)
    orig = ref_bn.inplace_abn
    ref_bn.inplace_abn = lambda *a: RefABN.forward(Ctx(), *a)
    m.train()
    assert m.finaldrop1.p == 0.5
    seen = {}
    hook = m.finaldrop1.register_forward_hook(
        lambda mod, inp, out: seen.update(keep=(out.detach().abs().sum(dim=(2, 3)) > 0), inp=inp[0].detach().clone(), out=out.detach().clone()))
    xt = torch.from_numpy(np.random.RandomState(16).standard_normal((4, 3, 64, 64)).astype(np.float32))
    torch.manual_seed(123)
    with torch.no_grad():
        yt = m(xt).numpy()
    hook.remove()
    ref_bn.inplace_abn = orig
    keep = seen["keep"]
    # Dropout2d semantics: kept channels are scaled by 1 / (1 - p), dropped ones are zero
    assert torch.allclose(seen["out"], seen["inp"] * (keep.float() * 2.0).view(4, 64, 1, 1))
    assert 0.3 < keep.float().mean().item() < 0.7
    np.savez_compressed(os.path.join(OUT, "linknet34_dropout.npz"), train_x=xt.numpy(), keep=keep.numpy(), train_logits=yt)


def unet_vectors():
    """UNet (lib/models/unet.py) and UNetABN (lib/models/unet_abn.py, InPlaceABN through the pure-torch shim of the
    un-vendored backend: parity unpinned at that boundary) in eval mode -> tests/golden/unet.npz."""
    from lib.models.unet import UNet
    from lib.models.unet_abn import UNetABN

    x = torch.from_numpy(np.random.RandomState(21).standard_normal((2, 3, 64, 96)).astype(np.float32))
    out = dict(x=x.numpy())
    for name, cls, abn in (("unet", UNet, False), ("unet_abn", UNetABN, True)):
        m = cls()
        res = m.load_state_dict(synth.unet_state_dict(seed=8, abn=abn), strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        m.eval()
        with torch.no_grad():
            out[name + "_logits"] = m(x).numpy()
    np.savez_compressed(os.path.join(OUT, "unet.npz"), **out)


def inplace_abn_vectors():
    """lib.modules.abn.functions.InPlaceABN (the reference's own autograd glue: running-statistics update, saved tensors,
    eval-mode shortcut) driven through the pure-torch stand-in for the un-vendored backend: training and eval mode,
    forward outputs, updated running statistics and all three gradients."""
    from lib.modules.abn.functions import InPlaceABN as RefABN

    class Ctx:
        """Stand-in for the autograd context: torch >= 2 rejects the reference's mark_dirty(x, running_mean,
        running_var) ("dirty tensors must be outputs"), so its forward / backward bodies are called directly."""

        def mark_dirty(self, *tensors):
            pass

        def save_for_backward(self, *tensors):
            self.saved_tensors = tensors

    rs = np.random.RandomState(21)
    out = {}
    x0 = rs.standard_normal((3, 6, 5, 8)).astype(np.float32) * 1.5 + 0.3
    w0 = (rs.uniform(0.5, 1.5, 6) * np.where(rs.rand(6) < 0.4, -1, 1)).astype(np.float32)
    b0 = (rs.standard_normal(6) * 0.2).astype(np.float32)
    rm0 = (rs.standard_normal(6) * 0.1).astype(np.float32)
    rv0 = rs.uniform(0.5, 1.5, 6).astype(np.float32)
    g0 = rs.standard_normal(x0.shape).astype(np.float32)
    out.update(x=x0, weight=w0, bias=b0, running_mean=rm0, running_var=rv0, grad=g0)
    for mode, training in (("train", True), ("eval", False)):
        w, b = torch.from_numpy(w0.copy()), torch.from_numpy(b0.copy())
        rm, rv = torch.from_numpy(rm0.copy()), torch.from_numpy(rv0.copy())
        ctx = Ctx()
        with torch.no_grad():
            z = RefABN.forward(ctx, torch.from_numpy(x0.copy()), w, b, rm, rv, training, 0.1, 1e-5, "leaky_relu", 0.01)
            zc = z.clone()
            grads = RefABN.backward(ctx, torch.from_numpy(g0.copy()))
        out[mode + "_z"] = zc.numpy()
        out[mode + "_running_mean"], out[mode + "_running_var"] = rm.numpy(), rv.numpy()
        out[mode + "_dx"], out[mode + "_dweight"], out[mode + "_dbias"] = [t.numpy() for t in grads[:3]]
    np.savez_compressed(os.path.join(OUT, "abn.npz"), **out)
    return dict(shape=list(x0.shape))


def predict_tiled_vector():
    """inria_submit.predict_tiled (:237-257) on CPU: same calls, without .cuda(); tile 64 / step 32, with and
    without D4 TTA, plus the submit threshold (:305)."""
    m = UNet16()
    m.load_state_dict(synth.vgg_unet_state_dict("unet16", seed=2))
    m.eval()
    image = synth.image_u8(9, 96, 80)
    t = aug.Sequential([aug.ImageOnly(aug.NormalizeImage(mean=INRIA_MEAN, std=INRIA_STD))])
    d = dict(image=image)
    for tta in (False, True):
        x, _ = t(image)
        slicer = ImageSlicer(x.shape, 64, 32, weight='pyramid')
        patches = slicer.split(x)
        if tta:
            patches = aug.tta_d4_aug(patches)
        ds = InMemoryDataset(patches, None)
        preds = []
        with torch.no_grad():
            for i in range(0, len(ds), 4):
                xb = torch.stack([ds[j] for j in range(i, min(i + 4, len(ds)))])
                y = torch.sigmoid(m(xb)).numpy()
                preds.extend(np.moveaxis(y, 1, -1))
        if tta:
            preds = aug.tta_d4_deaug(preds)
        mask = slicer.merge(preds, dtype=np.float32)
        key = "tta" if tta else "plain"
        d[key + "_tiles"] = np.stack(preds)
        d[key + "_merged"] = mask
        d[key + "_mask"] = ((mask > 0.5) * 255).astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "predict_tiled.npz"), **d)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    if len(sys.argv) > 1:                 # regenerate single files: python oracle/make_golden.py loss_extra_vectors ...
        for name in sys.argv[1:]:
            globals()[name]()
        return
    kats = dict(slicer=slicer_kats(), pyramid=weight_vectors(), loss=loss_vectors(), zf_unet_cfg1=zf_unet_vectors(), fcdensenet67=fcdensenet_vectors(), linknet34=linknet34_vectors(), inplace_abn=inplace_abn_vectors(),
                torch_version=torch.__version__, numpy_version=np.__version__)
    split_merge_vectors()
    normalize_vectors()
    tta_vectors()
    model_vectors()
    predict_tiled_vector()
    loss_extra_vectors()
    linknet34_dropout_vectors()
    unet_vectors()
    with open(os.path.join(OUT, "kats.json"), "w") as fh:
        json.dump(kats, fh, indent=1)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
