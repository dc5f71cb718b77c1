"""-m gpu: the callers and data formats either side of the hot path: device-side TiledImageDataset, the fused
train() / validate() loops, and the file-level submit pipeline with decode / write thread pools."""
import os

import numpy as np
import pytest
import torch

import snb_b200  # noqa: F401
from oracle import nets_oracle as no
from oracle import synth
from oracle import tiles_oracle as to
from snb_b200 import inria_submit as sub
from snb_b200 import torch_train as TT
from snb_b200.lib import augmentations as aug
from snb_b200.lib import losses, metrics
from snb_b200.lib.common import TiledImageDataset, cut_dataset_in_patches

pytestmark = pytest.mark.gpu


def test_tiled_image_dataset_matches_reference_semantics(cuda):
    """lib/common.py:116-159: item i = (normalised patch CHW float, mask 1HW long) of crop i; here cut + normalise + layout
    run on the device in one launch, bit-exact against the oracle's cut_patch + NormalizeImage + moveaxis + .float()."""
    image = synth.image_u8(3, 150, 170)
    mask = (synth.gt_mask_u8(4, 150, 170) * 255).astype(np.uint8)
    t = aug.Sequential([aug.ImageOnly(aug.NormalizeImage(mean=sub.INRIA_MEAN, std=sub.INRIA_STD))])
    ds = TiledImageDataset(image=image, mask=mask, tile_size=64, tile_step=0, transform=t)         # step 0 -> tile_size // 2
    so = to.SlicerOracle(image.shape, 64, 32)
    assert len(ds) == len(so.crops) and ds.slicer.tile_step == 32
    norm = to.normalize_image(image)
    for i in (0, 5, len(ds) - 1, -1):
        x, m = ds[i]
        assert x.dtype == torch.float32 and x.shape == (3, 64, 64) and m.dtype == torch.int64 and m.shape == (1, 64, 64)
        want_x = np.moveaxis(so.cut_patch(norm, i), -1, 0).astype(np.float32)
        assert np.array_equal(x.cpu().numpy(), want_x)
        assert np.array_equal(m.cpu().numpy()[0], so.cut_patch(mask, i))
    xb, mb = ds.batch(2, 7)
    assert xb.shape == (7, 3, 64, 64) and torch.equal(xb[3], ds[5][0]) and torch.equal(mb[3], ds[5][1])
    with pytest.raises(IndexError):
        ds[len(ds)]
    with pytest.raises(ValueError):
        TiledImageDataset(image=image, mask=mask[:-1], tile_size=64)
    # no transform: raw patches, moveaxis + .float() only
    raw = TiledImageDataset(image=image, mask=mask, tile_size=64, tile_step=32)
    x, _ = raw[4]
    assert np.array_equal(x.cpu().numpy(), np.moveaxis(so.cut_patch(image, 4), -1, 0).astype(np.float32))
    # Inria.cut_dataset_in_patches on the device
    for k, ti, tm in cut_dataset_in_patches([image], [mask], 64):
        tiles = so.split(image)
        assert ti.shape[0] == len(tiles) and np.array_equal(ti.cpu().numpy(), np.stack(tiles))
        assert np.array_equal(tm.cpu().numpy(), np.stack(so.split(mask)))


def test_fused_validate_and_train_loops(cuda):
    """torch_train.validate / train with one fused reduction per batch and one host copy per epoch: the meters equal the
    values the loss / metric modules return batch by batch, and a few epochs of train() lower the loss."""
    from snb_b200.lib.models import LinkNet34

    torch.manual_seed(1)
    m = LinkNet34(pretrained=False).cuda()
    rs = np.random.RandomState(9)
    yy, xx = np.mgrid[0:64, 0:64]
    target = ((yy // 16 + xx // 16) % 2)[None, None].repeat(4, 0).astype(np.int64)
    batches = [(torch.from_numpy(rs.standard_normal((4, 3, 64, 64)).astype(np.float32)), torch.from_numpy(target)) for _ in range(3)]
    crit = losses.BCEWithLogitsLossAndSmoothJaccard()
    mets = {'iou': metrics.JaccardScore(), 'acc': metrics.PixelAccuracy()}
    pr = metrics.PRCurveMeter()
    vl, vs = TT.validate(m, crit, batches, 0, mets, pr_meter=pr)
    assert vl.count == 3 and set(vs) == {'iou', 'acc'}
    want_loss, want_iou, want_acc = [], [], []
    m.eval()
    with torch.no_grad():
        for x, y in batches:
            out = m(x.cuda())
            want_loss.append(float(crit(out, y.cuda())))
            want_iou.append(float(mets['iou'](out, y.cuda())))
            want_acc.append(float(mets['acc'](out, y.cuda())))
    assert vl.avg == pytest.approx(np.mean(want_loss), rel=1e-5) and vl.val == pytest.approx(want_loss[-1], rel=1e-5)
    assert vs['iou'].avg == pytest.approx(np.mean(want_iou), rel=1e-5) and vs['acc'].avg == pytest.approx(np.mean(want_acc), rel=1e-6)
    assert int(pr.tp[0]) + int(pr.fn[0]) == int(target.sum())                  # PR counts of the last batch
    # against the CPU oracle's loss on the device logits of the last batch
    assert want_loss[-1] == pytest.approx(float(no.bce_jaccard(out.cpu(), batches[-1][1])), rel=1e-4)
    opt = torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9)
    hist = [TT.train(m, crit, opt, batches, e, mets)[0].avg for e in range(4)]
    assert all(np.isfinite(hist)) and hist[-1] < hist[0], hist
    # a criterion without fused coefficients takes the generic (autograd) path
    class Plain(torch.nn.Module):
        def forward(self, o, t):
            return torch.nn.functional.binary_cross_entropy_with_logits(o, t.float())
    pl, _ = TT.train(m, Plain(), opt, batches[:1], 0, {})
    assert np.isfinite(pl.avg)


def test_file_submitter_pipeline(cuda, tmp_path):
    """inria_submit.main's loop (inria_submit.py:291-306) with decode and imwrite thread pools around the streaming
    predictor: every written mask equals the mask of a direct predict_device call on the same image."""
    cv2 = pytest.importorskip("cv2")
    from snb_b200.lib.models import UNet16

    m = UNet16()
    m.load_state_dict(synth.vgg_unet_state_dict("unet16", seed=2))
    m = m.cuda().eval()
    shape = (96, 128, 3)
    paths = []
    for i in range(7):
        p = str(tmp_path / ("img_%02d.png" % i))
        cv2.imwrite(p, synth.image_u8(40 + i, shape[0], shape[1]))
        paths.append(p)
    pred = sub.TiledPredictor(m, shape, 64, 32, batch_size=4, tta=False)
    fs = sub.FileSubmitter(pred, decoders=3, writers=2)
    outs = fs.run(paths, str(tmp_path / "out"), suffix=".png")
    fs.close()
    assert len(outs) == 7 and all(os.path.exists(o) for o in outs)
    for p, o in zip(paths, outs):
        _, mask = pred.predict_device(torch.from_numpy(cv2.imread(p, cv2.IMREAD_COLOR)).cuda())
        got = cv2.imread(o, cv2.IMREAD_GRAYSCALE)
        assert np.array_equal(got, mask.cpu().numpy()[..., 0]), p
