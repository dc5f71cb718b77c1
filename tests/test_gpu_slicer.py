"""-m gpu: ImageSlicer split / merge kernels through the C ABI against the oracle and the reference's vectors."""
import os

import numpy as np
import pytest
import torch

import snb_b200  # noqa: F401
from oracle import tiles_oracle as to
from snb_b200 import _native as N
from snb_b200.lib.tiles import ImageSlicer

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["u8c3", "f32c1", "u8_2d", "f64c3", "tiny_multi_reflect"])
def test_split_matches_reference_vectors(cuda, golden_dir, name):
    g = np.load(os.path.join(golden_dir, "split_merge.npz"))
    img = g[name + "_image"]
    tile, step = (int(v) for v in g[name + "_cfg"])
    s = ImageSlicer(img.shape, tile, step)
    tiles = s.split(img)
    assert len(tiles) == g[name + "_tiles"].shape[0]
    assert tiles[0].shape == g[name + "_tiles"].shape[1:] and tiles[0].dtype == img.dtype
    assert np.array_equal(np.stack(tiles), g[name + "_tiles"])
    assert np.array_equal(s.cut_patch(img, min(3, len(s.crops) - 1)), g[name + "_cut3"])
    with pytest.raises(AssertionError):
        s.split(np.zeros((img.shape[0] + 1,) + img.shape[1:], img.dtype))


@pytest.mark.parametrize("shape,tile,step,dtype", [
    ((300, 260, 3), 128, 64, np.uint8), ((129, 77, 1), 32, 20, np.float32), ((64, 64, 3), 64, 64, np.float64),
    ((50, 70), 256, 128, np.uint8), ((33, 35, 2), 16, 9, np.int16)])
def test_split_matches_oracle(cuda, shape, tile, step, dtype):
    rs = np.random.RandomState(1)
    img = rs.randint(0, 255, shape).astype(dtype)
    want = np.stack(to.SlicerOracle(shape, tile, step).split(img))
    got = np.stack(ImageSlicer(shape, tile, step).split(img))
    assert got.dtype == want.dtype and np.array_equal(got, want)


def test_split_constant_border(cuda):
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(2)
    img = rs.randint(0, 255, (40, 50, 3)).astype(np.uint8)
    s = ImageSlicer(img.shape, 32, 24)
    padded = cv2.copyMakeBorder(img, s.margin_top, s.margin_bottom, s.margin_left, s.margin_right,
                                borderType=cv2.BORDER_CONSTANT, value=7)
    want = np.stack([padded[y:y + 32, x:x + 32] for x, y, _, _ in s.crops])
    got = np.stack(s.split(img, borderType=cv2.BORDER_CONSTANT, value=7))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("weight", ["mean", "pyramid"])
def test_merge_matches_reference_vectors(cuda, golden_dir, weight):
    g = np.load(os.path.join(golden_dir, "split_merge.npz"))
    s = ImageSlicer((37, 53, 3), 16, 8, weight=weight)
    out = s.merge(list(g["merge_%s_tiles" % weight]))
    assert out.dtype == np.float32 and np.array_equal(out, g["merge_%s_out" % weight])   # bit-exact (f64 path)
    ident = s.merge(s.split(g["merge_%s_identity_in" % weight]))
    assert np.array_equal(ident, g["merge_%s_identity_in" % weight])
    with pytest.raises(ValueError):
        s.merge(list(g["merge_%s_tiles" % weight])[:-1])


def test_merge_u8_vectors(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "split_merge.npz"))
    s = ImageSlicer((37, 53, 3), 16, 8, weight="mean")
    assert np.array_equal(s.merge(list(g["merge_u8_tiles"]), dtype=np.uint8), g["merge_u8_out"])


@pytest.mark.parametrize("shape,tile,step,weight", [((300, 260), 128, 64, "pyramid"), ((129, 77), 32, 20, "pyramid"),
                                                    ((96, 80), 64, 32, "mean"), ((50, 70), 256, 128, "pyramid")])
def test_merge_matches_oracle(cuda, shape, tile, step, weight):
    rs = np.random.RandomState(3)
    so = to.SlicerOracle(shape, tile, step, weight=weight)
    tiles = [rs.rand(tile, tile, 1).astype(np.float32) for _ in so.crops]
    want = so.merge(tiles)
    got = ImageSlicer(shape, tile, step, weight=weight).merge(tiles)
    assert np.array_equal(got, want)


def test_merge_tta_and_mask(cuda, golden_dir):
    """snb_merge with 8 D4 views per tile == tta_d4_deaug + merge + threshold of the reference."""
    g = np.load(os.path.join(golden_dir, "tta.npz"))
    rs = np.random.RandomState(4)
    so = to.SlicerOracle((40, 44), 16, 8, weight="pyramid")
    views = [rs.rand(16, 16, 1).astype(np.float32) for _ in range(8 * len(so.crops))]
    want = so.merge(to.tta_d4_deaug(views))
    s = ImageSlicer((40, 44), 16, 8, weight="pyramid")
    t = torch.from_numpy(np.stack(views)).cuda()
    out = torch.empty((40, 44, 1), dtype=torch.float32, device="cuda")
    mask = torch.empty((40, 44, 1), dtype=torch.uint8, device="cuda")
    N.check(N.lib().snb_merge(s.handle, N.ptr(t), N.DT_F32, 1, 8, N.ptr(s.weight_on_device()), N.ptr(out), N.DT_F32,
                              N.ptr(mask), 0.5, N.stream_ptr()))
    assert np.array_equal(out.cpu().numpy(), want)
    assert np.array_equal(mask.cpu().numpy(), ((want > 0.5) * 255).astype(np.uint8))
    # the de-augmentation alone against the reference's own vector
    deaug = np.stack(to.tta_d4_deaug(list(g["preds"])))
    assert np.array_equal(deaug, g["deaug"])


@pytest.mark.parametrize("tta", range(8))
def test_split_norm_layouts(cuda, tta):
    """Fused split: u8 -> LUT normalise -> D4 view -> NCHW fp32 (bit-exact) and PATCH32 bf16 rows."""
    img = np.random.RandomState(5).randint(0, 256, (70, 90, 3)).astype(np.uint8)
    T, step = 32, 24
    so = to.SlicerOracle(img.shape, T, step)
    norm = to.normalize_image(img)                        # float64, as the reference pipeline
    tiles = [to.d4_views(t)[tta] for t in so.split(norm)]
    want = to.to_nchw_float(tiles)                         # float32 [n][3][T][T]
    s = ImageSlicer(img.shape, T, step)
    n = len(s.crops)
    d_img = torch.from_numpy(img).cuda()
    lut = torch.from_numpy(to.normalize_lut()).cuda()
    out = torch.empty((n, 3, T, T), dtype=torch.float32, device="cuda")
    N.check(N.lib().snb_split_norm_u8(s.handle, N.ptr(d_img), 3, N.ptr(lut), tta, N.LAYOUT_NCHW_F32, N.ptr(out), 0, n,
                                      N.stream_ptr()))
    assert np.array_equal(out.cpu().numpy(), want)
    rows = torch.empty((n, T, T, 32), dtype=torch.bfloat16, device="cuda")
    N.check(N.lib().snb_split_norm_u8(s.handle, N.ptr(d_img), 3, N.ptr(lut), tta, N.LAYOUT_PATCH32, N.ptr(rows), 0, n,
                                      N.stream_ptr()))
    x = torch.from_numpy(want)
    cols = torch.nn.functional.unfold(x, 3, padding=1).reshape(n, 3, 9, T, T)      # [n][c][tap][y][x]
    want_rows = torch.zeros((n, T, T, 32))
    want_rows[..., :27] = cols.permute(0, 3, 4, 2, 1).reshape(n, T, T, 27)         # k = tap*3 + c
    assert torch.equal(rows.cpu().float(), want_rows.to(torch.bfloat16).float())
    # nn.Module entry: NCHW fp32 -> PATCH32
    rows2 = torch.empty_like(rows)
    N.check(N.lib().snb_nchw_f32_to_patch32(N.ptr(out), n, 3, T, T, N.ptr(rows2), 0, N.stream_ptr()))
    assert torch.equal(rows2, rows)
    # packed 3-channel bf16 tiles (6 bytes per pixel: the first-layer kernel builds its operand rows on chip)
    nhwc = torch.full((n, T, T, 3), float("nan"), dtype=torch.bfloat16, device="cuda")
    N.check(N.lib().snb_split_norm_u8(s.handle, N.ptr(d_img), 3, N.ptr(lut), tta, N.LAYOUT_NHWC3_BF16, N.ptr(nhwc), 0, n,
                                      N.stream_ptr()))
    assert torch.equal(nhwc.cpu().float(), x.permute(0, 2, 3, 1).to(torch.bfloat16).float())
    nhwc2 = torch.empty_like(nhwc)
    N.check(N.lib().snb_nchw_f32_to_nhwc3(N.ptr(out), n, T, T, N.ptr(nhwc2), N.stream_ptr()))
    assert torch.equal(nhwc2, nhwc)
    part = torch.zeros((2, T, T, 3), dtype=torch.bfloat16, device="cuda")           # a tile sub-range
    N.check(N.lib().snb_split_norm_u8(s.handle, N.ptr(d_img), 3, N.ptr(lut), tta, N.LAYOUT_NHWC3_BF16, N.ptr(part), n - 2, 2,
                                      N.stream_ptr()))
    assert torch.equal(part, nhwc[n - 2:])


def test_full_size_roundtrip_properties(cuda):
    """BASELINE size (5000 x 5000, 512/384): plan KATs, split->merge identity, constant-in -> constant-out."""
    s = ImageSlicer((5000, 5000, 1), 512, 384, weight="pyramid")
    assert len(s.crops) == 169 and (s.margin_left, s.margin_right, s.margin_top, s.margin_bottom) == (60,) * 4
    g = torch.Generator(device="cuda").manual_seed(0)
    img = torch.rand((5000, 5000, 1), device="cuda", generator=g)
    tiles = torch.stack(s.split(img))
    assert tiles.shape == (169, 512, 512, 1)
    assert torch.equal(tiles[0, 60:, 60:, 0], img[:452, :452, 0])                  # interior copy
    assert torch.equal(tiles[0, 59, 60:, 0], img[1, :452, 0])                      # reflect-101 row
    assert torch.equal(tiles[168, :452, :452, 0], img[4548:, 4548:, 0])
    merged = s.merge(tiles)
    assert torch.equal(merged, img)                                                # weighted mean of equal values
    ones = s.merge(torch.full((169, 512, 512, 1), 0.25, device="cuda"))
    assert torch.equal(ones, torch.full_like(ones, 0.25))


@pytest.mark.parametrize("shape,tile,step,weight", [
    ((200, 264), 64, 32, "pyramid"),     # tile = 2 * step: every pixel sees the crop to its left
    ((120, 200), 64, 48, "pyramid"),     # narrow overlap band (16 px), margins 4
    ((100, 64), 64, 64, "pyramid"),      # no overlap, one crop column: pure copy path
    ((50, 72), 128, 64, "pyramid"),      # single crop, margins 28 / 39
    ((96, 80), 64, 32, "mean"),
    ((333, 520), 128, 96, "pyramid"),    # odd height, several periods
])
def test_merge_periodic_kernel_bit_exact(cuda, shape, tile, step, weight):
    """The float32 one-channel fast path (merge_f32c1_period_kernel: loop-invariant norm, FMA-corrected division by an
    invariant, copy of single-cover pixels) against the numpy restatement of lib/tiles.py:137-161, bit for bit, with
    values that stress the division (zeros, float32 denormals, 1e30) and the threshold output."""
    rs = np.random.RandomState(11)
    so = to.SlicerOracle(shape, tile, step, weight=weight)
    s = ImageSlicer(shape, tile, step, weight=weight)
    assert s.margin_left % 4 == 0 and shape[1] % 4 == 0           # geometry of the fast path
    tiles = rs.rand(len(so.crops), tile, tile, 1).astype(np.float32)
    special = rs.rand(*tiles.shape)
    tiles[special < 0.05] = 0.0
    tiles[(special > 0.05) & (special < 0.07)] = 1e-42
    tiles[(special > 0.07) & (special < 0.08)] = 1e30
    want = so.merge(list(tiles))
    t = torch.from_numpy(tiles).cuda()
    out = torch.full(shape + (1,), float("nan"), dtype=torch.float32, device="cuda")
    mask = torch.full(shape + (1,), 7, dtype=torch.uint8, device="cuda")
    N.check(N.lib().snb_merge(s.handle, N.ptr(t), N.DT_F32, 1, 1, N.ptr(s.weight_on_device()), N.ptr(out), N.DT_F32,
                              N.ptr(mask), 0.5, N.stream_ptr()))
    got = out.cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(mask.cpu().numpy(), ((want > 0.5) * 255).astype(np.uint8))


def test_merge_full_size_random_rows_against_oracle(cuda):
    """BASELINE size: 169 random probability tiles; rows sampled across all crop-row patterns are compared bit for bit
    with the float64 numpy accumulation of lib/tiles.py:146-161 restricted to those rows."""
    s = ImageSlicer((5000, 5000, 1), 512, 384, weight="pyramid")
    g = torch.Generator(device="cuda").manual_seed(5)
    tiles = torch.rand((169, 512, 512, 1), device="cuda", generator=g)
    merged = s.merge(tiles)[..., 0].cpu().numpy()
    w = s.weight_on_device().cpu().numpy()
    rows = [0, 1, 67, 68, 323, 324, 325, 451, 452, 2500, 4547, 4548, 4931, 4932, 4999]
    th = tiles[..., 0].cpu().numpy()
    for y in rows:
        Y = y + 60
        acc, nrm = np.zeros(5120, np.float64), np.zeros(5120, np.float64)
        for iy in range(13):
            ty = Y - iy * 384
            if 0 <= ty < 512:
                for ix in range(13):
                    acc[ix * 384:ix * 384 + 512] += th[iy * 13 + ix, ty].astype(np.float64) * w[ty]
                    nrm[ix * 384:ix * 384 + 512] += w[ty]
        want = (acc / np.clip(nrm, np.finfo(np.float64).eps, None)).astype(np.float32)[60:5060]
        assert np.array_equal(merged[y].view(np.uint32), want.view(np.uint32)), y
