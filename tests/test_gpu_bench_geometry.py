"""-m gpu: parity at the geometry bench.py times (VERDICT r1 "weak" item 1).

The unit tests run small shapes where `snb_conv_create` picks other kernels than at the benchmarked shapes
(wave-aware N = 128 tiles on the 32x32 / 16x16 layers, resident weights, conv_halo_kernel<256,...>).  Here the plans are
built exactly as bench.py builds them -- default SNB_CONV_MODE, UNet16 at 13 and 85 x 512 x 512, FCDenseNet67 / ZF_UNET at
121 / 176 x 224 x 224 (and at 44, round 1's batch) -- and checked against the CPU oracle (fp32, and with bf16-rounded conv operands) on a few tiles of the
batch, then through the whole tiled pipeline at 512 / 384 on an image small enough for the oracle.
"""
import numpy as np
import pytest
import torch

import snb_b200  # noqa: F401
from oracle import nets_oracle as no
from oracle import synth
from oracle import tiles_oracle as to
from snb_b200 import inria_submit as sub

pytestmark = pytest.mark.gpu

BF16_PROB_TOL = 2e-2


@pytest.fixture(autouse=True)
def default_conv_mode(monkeypatch):
    for k in ("SNB_CONV_MODE", "SNB_FUSE_POOL", "SNB_FUSE_PRE", "SNB_SCATTER", "SNB_A_STAGES", "SNB_B_STAGES"):
        monkeypatch.delenv(k, raising=False)


def _check_tiles(probs, x, picks, fwd, fwd_q):
    """probs: device probabilities [n, h, w]; oracle on the picked batch entries only (a 512x512 tile costs ~1 s)."""
    worst_p, worst_l = 0.0, 0.0
    for i in picks:
        xi = x[i:i + 1].cpu()
        with torch.no_grad():
            ref = fwd(xi)[0, 0]
            q = fwd_q(xi)[0, 0]
        got = probs[i].cpu()
        worst_p = max(worst_p, (got - torch.sigmoid(ref)).abs().max().item())
        # against the oracle evaluated with bf16-rounded conv operands the logits agree much more tightly
        got_logit = torch.logit(got.double().clamp(1e-12, 1 - 1e-12)).float()
        sel = q.abs() < 8        # sigmoid saturates in float32 beyond that
        worst_l = max(worst_l, ((got_logit - q).abs()[sel].max() / max(1.0, q.abs().max().item())).item())
    return worst_p, worst_l


@pytest.mark.parametrize("nb", [13, 85])
def test_unet16_plan_at_bench_geometry(cuda, nb):
    """UNet16, 85 tiles of 512x512 per launch (the plan bench.py replays; 13 = round 1's), default kernel selection."""
    from snb_b200.lib.models import UNet16

    sd = synth.vgg_unet_state_dict("unet16", seed=0)           # bench.py's weights
    m = UNet16()
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((nb, 3, 512, 512), device="cuda", generator=g)
    plan = m.plan(nb, 512, 512, sigmoid=True)
    plan.load_nchw(x)
    probs = plan.run().clone()
    torch.cuda.synchronize()
    assert probs.shape == (nb, 512, 512) and torch.isfinite(probs).all()
    p_err, l_err = _check_tiles(probs, x, (0, nb // 2, nb - 1), lambda t: no.unet_vgg_forward(sd, t, "unet16"),
                                lambda t: no.unet_vgg_forward(sd, t, "unet16", quant=no.bf16_round))
    print("unet16 %dx512x512: max |p - oracle| = %.3g, logits vs bf16 oracle %.3g" % (nb, p_err, l_err))
    assert p_err < BF16_PROB_TOL, p_err
    assert l_err < 0.02, l_err
    # the nn.Module path (logits, no sigmoid) at the same geometry agrees with the plan
    with torch.no_grad():
        y = m(x[:2])
    assert (torch.sigmoid(y[:, 0]) - probs[:2]).abs().max().item() < 2e-3


@pytest.mark.parametrize("arch,nb", [("fcdensenet67", 44), ("zf_unet", 44), ("fcdensenet67", 121), ("zf_unet", 176)])
def test_224_plans_at_bench_batch(cuda, arch, nb):
    """configs[4] (FCDenseNet67) and ZF_UNET at the tile batches bench.py uses for them (121 / 176 x 224 x 224; 44 in round 1)."""
    from snb_b200.lib.models import FCDenseNet67, ZF_UNET

    if arch == "fcdensenet67":
        sd = synth.fcdensenet_state_dict(seed=0)
        m = FCDenseNet67(n_classes=1)
        fwd = lambda t: no.fcdensenet_forward(sd, t)
        fwd_q = lambda t: no.fcdensenet_forward(sd, t, quant=no.bf16_round)
    else:
        sd = synth.zf_unet_state_dict(seed=0)
        m = ZF_UNET()
        fwd = lambda t: no.zf_unet_forward(sd, t)
        fwd_q = lambda t: no.zf_unet_forward(sd, t, quant=no.bf16_round, fold=True)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((nb, 3, 224, 224), device="cuda", generator=g)
    plan = m.plan(nb, 224, 224, sigmoid=True)
    plan.load_nchw(x)
    probs = plan.run().clone()
    torch.cuda.synchronize()
    p_err, l_err = _check_tiles(probs, x, (0, nb // 2, nb - 1), fwd, fwd_q)
    print("%s %dx224x224: max |p - oracle| = %.3g, logits vs bf16 oracle %.3g" % (arch, nb, p_err, l_err))
    assert p_err < BF16_PROB_TOL, p_err
    assert l_err < 0.03, l_err


def test_predict_tiled_512_384_against_oracle_pipeline(cuda):
    """The headline pipeline (split -> UNet16 -> pyramid merge -> threshold) at tile 512 / step 384 on a 1024 x 1408
    image (12 crops, margins and overlaps as in the 5000 x 5000 case) against the oracle's normalise / split / forward /
    merge; mask bytes may differ only where the oracle's probability is within the tolerance of 0.5."""
    from snb_b200.lib.models import UNet16

    sd = synth.vgg_unet_state_dict("unet16", seed=0)
    m = UNet16()
    m.load_state_dict(sd)
    m = m.cuda().eval()
    image = synth.image_u8(21, 1024, 1408)
    pred = sub.TiledPredictor(m, image.shape, 512, 384, batch_size=13, tta=False)
    merged, mask = pred.predict_device(torch.from_numpy(image).cuda())
    merged, mask = merged.cpu().numpy(), mask.cpu().numpy()
    merged2, mask2 = pred.predict_device(torch.from_numpy(image).cuda())        # graph replay
    assert np.array_equal(merged2.cpu().numpy(), merged) and np.array_equal(mask2.cpu().numpy(), mask)

    # the tile batching is not allowed to change a single bit (kernel selection depends on it: N tile, waves, issuers):
    # 5 tiles per launch (3 launches, 3 padded slots) against all 12 in one launch
    pred5 = sub.TiledPredictor(m, image.shape, 512, 384, batch_size=5, tta=False)
    merged5, mask5 = pred5.predict_device(torch.from_numpy(image).cuda())
    assert np.array_equal(merged5.cpu().numpy(), merged) and np.array_equal(mask5.cpu().numpy(), mask)

    x = to.normalize_image(image)
    s = to.SlicerOracle(x.shape, 512, 384, weight="pyramid")
    assert len(s.crops) == pred.n_tiles
    tiles = to.to_nchw_float(s.split(x))
    probs = []
    with torch.no_grad():
        for i in range(0, len(tiles), 2):
            probs.append(torch.sigmoid(no.unet_vgg_forward(sd, torch.from_numpy(tiles[i:i + 2]), "unet16")).numpy())
    probs = np.concatenate(probs)
    want = s.merge(list(np.moveaxis(probs, 1, -1)), dtype=np.float32)
    err = np.abs(merged - want).max()
    print("predict_tiled 1024x1408 @512/384: max |p - oracle| = %.3g over %d crops" % (err, pred.n_tiles))
    assert err < BF16_PROB_TOL, err
    want_mask = ((want > 0.5) * 255).astype(np.uint8)
    flips = mask != want_mask
    assert np.all(np.abs(want[flips] - 0.5) < BF16_PROB_TOL)
    # feeding the ORACLE's probability tiles through the device merge reproduces the oracle's bytes (merge is bit-exact)
    pred.probs[:, 0, :, :, 0].copy_(torch.from_numpy(probs[:, 0]).cuda())
    m2, k2 = pred.merge_probs()
    assert np.array_equal(m2.cpu().numpy(), want) and np.array_equal(k2.cpu().numpy(), want_mask)
