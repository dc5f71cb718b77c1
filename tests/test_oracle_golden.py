"""The CPU oracle against the vectors produced by the reference itself (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import nets_oracle as no
from oracle import synth
from oracle import tiles_oracle as to


def test_slicer_plan_kats(kats):
    for c in kats["slicer"]["cases"]:
        s = to.SlicerOracle(tuple(c["shape"]), c["tile"], c["step"], image_margin=c["margin"])
        assert [s.margin_left, s.margin_right, s.margin_top, s.margin_bottom] == c["margins"]
        assert len(s.crops) == c["n_crops"]
        assert [list(x) for x in s.crops[:3]] == c["crops_head"]
        assert [list(x) for x in s.crops[-3:]] == c["crops_tail"]
        assert [sum(x[0] for x in s.crops), sum(x[1] for x in s.crops)] == c["crops_sum"]


def test_slicer_plan_errors(kats):
    for c in kats["slicer"]["errors"]:
        if c["error"] is None:
            to.SlicerOracle(tuple(c["shape"]), c["tile"], c["step"], image_margin=c["margin"])
        else:
            with pytest.raises(ValueError):
                to.SlicerOracle(tuple(c["shape"]), c["tile"], c["step"], image_margin=c["margin"])


@pytest.mark.parametrize("name", ["u8c3", "f32c1", "u8_2d", "f64c3", "tiny_multi_reflect"])
def test_split_bit_exact(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "split_merge.npz"))
    img = g[name + "_image"]
    tile, step = (int(v) for v in g[name + "_cfg"])
    s = to.SlicerOracle(img.shape, tile, step)
    tiles = np.stack(s.split(img))
    assert tiles.dtype == g[name + "_tiles"].dtype and tiles.shape == g[name + "_tiles"].shape
    assert np.array_equal(tiles, g[name + "_tiles"])
    assert np.array_equal(s.cut_patch(img, min(3, len(s.crops) - 1)), g[name + "_cut3"])


@pytest.mark.parametrize("weight", ["mean", "pyramid"])
def test_merge_bit_exact(golden_dir, weight):
    g = np.load(os.path.join(golden_dir, "split_merge.npz"))
    s = to.SlicerOracle((37, 53, 3), 16, 8, weight=weight)
    out = s.merge(list(g["merge_%s_tiles" % weight]))
    assert out.dtype == np.float32 and np.array_equal(out, g["merge_%s_out" % weight])
    ident = s.merge(s.split(g["merge_%s_identity_in" % weight]))
    assert np.array_equal(ident, g["merge_%s_identity_out" % weight])
    assert np.array_equal(ident, g["merge_%s_identity_in" % weight])  # weighted mean of equal values is exact


def test_merge_u8(golden_dir):
    g = np.load(os.path.join(golden_dir, "split_merge.npz"))
    s = to.SlicerOracle((37, 53, 3), 16, 8, weight="mean")
    assert np.array_equal(s.merge(list(g["merge_u8_tiles"]), dtype=np.uint8), g["merge_u8_out"])


def test_merge_rejects_wrong_count():
    s = to.SlicerOracle((37, 53, 3), 16, 8)
    with pytest.raises(ValueError):
        s.merge([np.zeros((16, 16, 1), np.float32)])


def test_pyramid_weight(golden_dir, kats):
    g = np.load(os.path.join(golden_dir, "pyramid.npz"))
    for n in (16, 24, 64):
        assert np.array_equal(to.pyramid_weight(n, n), g["w%d" % n])
    assert np.array_equal(to.pyramid_weight_loop(16, 16), g["w16"])
    w = to.pyramid_weight(512, 512)
    k = kats["pyramid"]["n512"]
    assert w.min() == k["min"] and w.max() == k["max"] and w.sum() == k["sum"]
    assert np.array_equal(w, w.T) and np.array_equal(w, w[::-1]) and np.array_equal(w, w[:, ::-1])


def test_normalize_lut(golden_dir):
    g = np.load(os.path.join(golden_dir, "normalize.npz"))
    assert np.array_equal(to.normalize_image(g["levels"]), g["out64"])
    lut = to.normalize_lut()
    assert lut.shape == (3, 256) and lut.dtype == np.float32
    assert np.array_equal(lut, g["chw_f32"][:, :, 0])   # [C][256][1] -> [C][256]


def test_tta(golden_dir):
    g = np.load(os.path.join(golden_dir, "tta.npz"))
    views = np.stack(to.tta_d4_aug(list(g["tiles"])))
    assert np.array_equal(views, g["views"])
    deaug = np.stack(to.tta_d4_deaug(list(g["preds"])))
    assert deaug.dtype == np.float32 and np.array_equal(deaug, g["deaug"])


@pytest.mark.parametrize("seed", [0, 3])
def test_loss_and_metrics(kats, seed):
    k = kats["loss"]["seed%d" % seed]
    logits, targets = synth.logits_targets(seed, tuple(k["shape"]))
    assert float(no.bce_jaccard(logits, targets)) == pytest.approx(k["bce_jaccard"], rel=1e-6)
    assert float(no.bce_with_sigmoid(logits, targets)) == pytest.approx(k["bce"], rel=1e-6)
    assert float(no.smooth_jaccard(logits, targets)) == pytest.approx(k["smooth_jaccard"], rel=1e-6)
    assert float(no.jaccard_score(logits, targets)) == pytest.approx(k["jaccard_score"], rel=1e-6)
    assert float(no.pixel_accuracy(logits, targets)) == pytest.approx(k["pixel_accuracy"], rel=1e-7)
    assert no.confusion_counts(torch.sigmoid(logits), targets).tolist() == k["counts"]
    tp, tn, fp, fn = no.pr_curve_counts(logits, targets)
    assert tp.tolist() == k["pr_tp"] and tn.tolist() == k["pr_tn"]
    assert fp.tolist() == k["pr_fp"] and fn.tolist() == k["pr_fn"]


@pytest.mark.parametrize("arch", ["unet16", "unet11"])
def test_model_logits(golden_dir, arch):
    g = np.load(os.path.join(golden_dir, "models.npz"))
    sd = synth.vgg_unet_state_dict(arch, seed=1)
    with torch.no_grad():
        y = no.unet_vgg_forward(sd, torch.from_numpy(g[arch + "_x"]), arch).numpy()
    assert y.shape == g[arch + "_logits"].shape
    assert np.abs(y - g[arch + "_logits"]).max() < 1e-5   # same ATen CPU kernels; order of ops identical


@pytest.mark.parametrize("tta", [False, True])
def test_predict_tiled_pipeline(golden_dir, tta):
    """inria_submit.predict_tiled restated with oracle parts only reproduces the reference's merged mask."""
    g = np.load(os.path.join(golden_dir, "predict_tiled.npz"))
    key = "tta" if tta else "plain"
    sd = synth.vgg_unet_state_dict("unet16", seed=2)
    image = g["image"]
    assert np.array_equal(image, synth.image_u8(9, 96, 80))
    x = to.normalize_image(image)
    s = to.SlicerOracle(x.shape, 64, 32, weight="pyramid")
    patches = s.split(x)
    if tta:
        patches = to.tta_d4_aug(patches)
    with torch.no_grad():
        probs = torch.sigmoid(no.unet_vgg_forward(sd, torch.from_numpy(to.to_nchw_float(patches)), "unet16")).numpy()
    preds = list(np.moveaxis(probs, 1, -1))
    if tta:
        preds = to.tta_d4_deaug(preds)
    assert np.abs(np.stack(preds) - g[key + "_tiles"]).max() < 2e-6
    merged = s.merge(preds, dtype=np.float32)
    assert np.abs(merged - g[key + "_merged"]).max() < 2e-6
    # merge itself is bit-exact when fed the reference's own tiles
    assert np.array_equal(s.merge(list(g[key + "_tiles"]), dtype=np.float32), g[key + "_merged"])
    assert np.array_equal(((g[key + "_merged"] > 0.5) * 255).astype(np.uint8), g[key + "_mask"])


def test_zf_unet_logits_and_config1(golden_dir, kats):
    """BASELINE configs[0]: ZF_UNET forward (eval, randomised BatchNorm buffers) + bce_jaccard + IoU + accuracy."""
    g = np.load(os.path.join(golden_dir, "zf_unet.npz"))
    sd = synth.zf_unet_state_dict(seed=4)
    with torch.no_grad():
        y = no.zf_unet_forward(sd, torch.from_numpy(g["small_x"])).numpy()
        y_fold = no.zf_unet_forward(sd, torch.from_numpy(g["small_x"]), fold=True).numpy()
    assert np.abs(y - g["small_logits"]).max() < 1e-4
    assert np.abs(y_fold - g["small_logits"]).max() < 1e-3       # folding BatchNorm into the conv is the same function
    k = kats["zf_unet_cfg1"]
    x224 = torch.from_numpy(np.random.RandomState(8).standard_normal((8, 3, 224, 224)).astype(np.float32))
    _, targets = synth.logits_targets(8, (8, 1, 224, 224))
    with torch.no_grad():
        logits = no.zf_unet_forward(sd, x224)
    assert np.abs(logits[:, :, ::7, ::7].numpy() - g["cfg1_logits_sample"]).max() < 2e-4
    assert float(no.bce_jaccard(logits, targets)) == pytest.approx(k["bce_jaccard"], rel=1e-5)
    assert float(no.jaccard_score(logits, targets)) == pytest.approx(k["jaccard_score"], rel=1e-4)
    assert float(no.pixel_accuracy(logits, targets)) == pytest.approx(k["pixel_accuracy"], abs=1e-5)


def test_fcdensenet67_logits(golden_dir, kats):
    g = np.load(os.path.join(golden_dir, "fcdensenet67.npz"))
    sd = synth.fcdensenet_state_dict(seed=5)
    assert len(sd) == 434 and kats["fcdensenet67"]["params"] == 3460353
    with torch.no_grad():
        y = no.fcdensenet_forward(sd, torch.from_numpy(g["x"])).numpy()
    assert np.abs(y - g["logits"]).max() < 1e-4


def test_linknet34_logits(golden_dir, kats):
    """BASELINE configs[1] model in eval mode; the reference ran with a pure-torch stand-in for the un-vendored
    `inplace_abn` backend (oracle/make_golden.py), so this row is "parity unpinned" at that boundary."""
    g = np.load(os.path.join(golden_dir, "linknet34.npz"))
    sd = synth.linknet34_state_dict(seed=6)
    assert len(sd) == kats["linknet34"]["n_keys"] == 294 and kats["linknet34"]["params"] == 21794721
    with torch.no_grad():
        y = no.linknet34_forward(sd, torch.from_numpy(g["x"])).numpy()
    assert y.shape == g["logits"].shape == (2, 1, 64, 96)
    assert np.abs(y - g["logits"]).max() < 1e-4


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_inplace_abn_restatement(golden_dir, mode):
    """oracle InPlaceABN forward / backward against the reference's own functions.py bodies (driven through the pure-torch
    stand-in for the un-vendored `inplace_abn` backend: parity unpinned at that boundary, pinned for the glue)."""
    g = np.load(os.path.join(golden_dir, "abn.npz"))
    t = lambda k: torch.from_numpy(g[k])
    training = mode == "train"
    z, var, rm, rv = no.inplace_abn_forward(t("x"), t("weight"), t("bias"), t("running_mean"), t("running_var"), training)
    assert np.abs(z.numpy() - g[mode + "_z"]).max() < 1e-5
    assert np.abs(rm.numpy() - g[mode + "_running_mean"]).max() < 1e-6
    assert np.abs(rv.numpy() - g[mode + "_running_var"]).max() < 1e-6
    dx, dw, db = no.inplace_abn_backward(z, t("grad"), var, t("weight"), t("bias"), training)
    assert np.abs(dx.numpy() - g[mode + "_dx"]).max() < 1e-5
    assert np.abs(dw.numpy() - g[mode + "_dweight"]).max() < 2e-4
    assert np.abs(db.numpy() - g[mode + "_dbias"]).max() < 2e-4
    if not training:
        assert not g["eval_dweight"].any() and not g["eval_dbias"].any()    # the reference's eval-mode shortcut


@pytest.mark.parametrize("seed,shape", [(0, (8, 1, 224, 224)), (3, (2, 1, 33, 17))])
def test_loss_gradients(golden_dir, seed, shape):
    """d loss / d logits of the oracle restatements under autograd == the reference modules' (lib/losses.py:31-75)."""
    g = np.load(os.path.join(golden_dir, "loss_grad.npz"))
    logits, targets = synth.logits_targets(seed, shape)
    for name, fn in (("bce_jaccard", no.bce_jaccard), ("smooth_jaccard", no.smooth_jaccard), ("bce", no.bce_with_sigmoid)):
        x = logits.clone().requires_grad_(True)
        (fn(x, targets) * 3.0).backward()
        got = x.grad.numpy().reshape(-1)
        want = g["seed%d_%s" % (seed, name)]
        got = got if got.size < 5000 else got[::97]
        assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()


def test_linknet34_train_mode_forward(golden_dir):
    """train() mode (batch statistics, Dropout2d off): oracle restatement against the reference module run with its own
    InPlaceABN.forward body (tests/golden/linknet34.npz train_*), logits and updated running statistics."""
    g = np.load(os.path.join(golden_dir, "linknet34.npz"))
    sd = synth.linknet34_state_dict(seed=6)
    with torch.no_grad():
        y, new = no.linknet34_forward_train(sd, torch.from_numpy(g["train_x"]))
    assert np.abs(y.numpy() - g["train_logits"]).max() < 1e-5
    for k in g.files:
        if "__" in k:
            assert np.abs(new[k.replace("__", ".")].numpy() - g[k]).max() < 1e-6, k


@pytest.mark.parametrize("seed,shape", [(0, (8, 1, 224, 224)), (3, (2, 1, 33, 17))])
def test_extra_losses(golden_dir, seed, shape):
    """JaccardLoss, FocalLossBinary, BCEWithSigmoidLoss(reduce=False / sum) restatements against the reference modules
    (lib/losses.py:18-28,46-53,78-101): values and gradients (tests/golden/loss_extra.npz)."""
    g = np.load(os.path.join(golden_dir, "loss_extra.npz"))
    logits, targets = synth.logits_targets(seed, shape)
    sample = lambda a: a if a.size < 5000 else a[::97]
    up = torch.from_numpy(np.random.RandomState(40 + seed).standard_normal(shape).astype(np.float32))
    cases = [("jaccard", no.jaccard_loss, None),
             ("focal_g2_mean", lambda x, t: no.focal_loss_binary(x, t, 2, True), None),
             ("focal_g1.5_sum", lambda x, t: no.focal_loss_binary(x, t, 1.5, False), None),
             ("focal_g0_mean", lambda x, t: no.focal_loss_binary(x, t, 0, True), None),
             ("bce_sum", lambda x, t: no.bce_elements(x, t).sum(), None),
             ("bce_elem", no.bce_elements, up)]
    for tag, fn, upstream in cases:
        x = logits.clone().requires_grad_(True)
        y = fn(x, targets)
        if upstream is None:
            assert float(y) == pytest.approx(float(g["seed%d_%s" % (seed, tag)]), rel=1e-6), tag
            (y * 3.0).backward()
        else:
            assert np.abs(sample(y.detach().numpy().reshape(-1)) - g["seed%d_%s" % (seed, tag)]).max() < 1e-6
            (y * upstream).sum().backward()
        want = g["seed%d_%s_grad" % (seed, tag)]
        assert np.abs(sample(x.grad.numpy().reshape(-1)) - want).max() <= 2e-6 * np.abs(want).max() + 1e-12, tag


def test_linknet34_train_mode_with_dropout(golden_dir):
    """Default Dropout2d(p=0.5) active: the restatement given the keep mask torch drew inside the reference module
    reproduces the reference's train-mode logits (tests/golden/linknet34_dropout.npz)."""
    g = np.load(os.path.join(golden_dir, "linknet34_dropout.npz"))
    sd = synth.linknet34_state_dict(seed=6)
    with torch.no_grad():
        y, _ = no.linknet34_forward_train(sd, torch.from_numpy(g["train_x"]), keep=torch.from_numpy(g["keep"]))
    assert np.abs(y.numpy() - g["train_logits"]).max() < 1e-5


@pytest.mark.parametrize("name,abn", [("unet", False), ("unet_abn", True)])
def test_unet_logits(golden_dir, name, abn):
    """UNet / UNetABN restatement against the reference modules in eval mode (tests/golden/unet.npz)."""
    g = np.load(os.path.join(golden_dir, "unet.npz"))
    sd = synth.unet_state_dict(seed=8, abn=abn)
    with torch.no_grad():
        y = no.unet_forward(sd, torch.from_numpy(g["x"]), abn=abn).numpy()
    assert y.shape == g[name + "_logits"].shape == (2, 1, 64, 96)
    assert np.abs(y - g[name + "_logits"]).max() < 2e-5
