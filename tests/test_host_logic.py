"""CPU: host-side mirror of the reference interface (ImageSlicer attributes, weight packing, model key names,
transforms) and the world_size-2 sharding / exchange logic over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import snb_b200  # noqa: F401
from oracle import nets_oracle as no
from oracle import synth
from oracle import tiles_oracle as to
from snb_b200 import dist as sdist
from snb_b200 import engine as E
from snb_b200.lib import augmentations as aug
from snb_b200.lib.models import FCDenseNet67, UNet11, UNet16, ZF_UNET
from snb_b200.lib.tiles import ImageSlicer, compute_patch_weight_loss


def test_image_slicer_attributes_match_reference(kats):
    for c in kats["slicer"]["cases"]:
        s = ImageSlicer(tuple(c["shape"]), c["tile"], c["step"], image_margin=c["margin"])
        assert (s.image_height, s.image_width, s.tile_size, s.tile_step) == (c["shape"][0], c["shape"][1], c["tile"], c["step"])
        assert [s.margin_left, s.margin_right, s.margin_top, s.margin_bottom] == c["margins"]
        assert len(s.crops) == c["n_crops"] and [list(v) for v in s.crops[:3]] == c["crops_head"]
        o = to.SlicerOracle(tuple(c["shape"]), c["tile"], c["step"], image_margin=c["margin"])
        assert s.crops == o.crops


def test_image_slicer_errors(kats):
    for c in kats["slicer"]["errors"]:
        if c["error"]:
            with pytest.raises(ValueError):
                ImageSlicer(tuple(c["shape"]), c["tile"], c["step"], image_margin=c["margin"])
    with pytest.raises(ValueError):
        ImageSlicer((64, 64), 32)                       # the default tile_step=0 raises in the reference too
    with pytest.raises(KeyError):
        ImageSlicer((64, 64), 32, 16, weight="gauss")
    s = ImageSlicer((64, 64), 32, 16)
    with pytest.raises(ValueError):
        s.merge([np.zeros((32, 32, 1), np.float32)])    # wrong tile count (lib/tiles.py:138-139)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_for_compute():
    s = ImageSlicer((64, 64), 32, 16)
    with pytest.raises(RuntimeError):
        s.split(np.zeros((64, 64), np.uint8))
    with pytest.raises(RuntimeError):
        UNet16()(torch.zeros(1, 3, 32, 32))


def test_pyramid_weight_is_the_reference_expression(golden_dir):
    g = np.load(os.path.join(golden_dir, "pyramid.npz"))
    for n in (16, 24, 64):
        assert np.array_equal(compute_patch_weight_loss(n, n)[0], g["w%d" % n])
    assert np.array_equal(ImageSlicer((64, 64), 16, 8, weight="pyramid").compute_weight(16), g["w16"])
    w = ImageSlicer((64, 64), 16, 8, weight="mean").compute_weight(16)
    assert w.dtype == np.float32 and np.all(w == 1)


def test_normalize_lut_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "normalize.npz"))
    n = aug.NormalizeImage(mean=to.INRIA_MEAN, std=to.INRIA_STD)
    assert np.array_equal(n(g["levels"]), g["out64"])
    assert np.array_equal(n.lut(), g["chw_f32"][:, :, 0])
    t = aug.Sequential([aug.ImageOnly(n)])
    assert aug.find_normalize(t) is n and aug.find_normalize(object()) is None


def _tap_list_conv(x_nhwc, packed, taps):
    """CPU emulation of what the kernel computes for one phase: sum over taps of shifted input @ W[tap]^T."""
    n, h, w, cin = x_nhwc.shape
    out = torch.zeros((n, h, w, packed.shape[1]))
    xp = F.pad(x_nhwc, (0, 0, 2, 2, 2, 2))
    for t, (dy, dx) in enumerate(taps):
        out += xp[:, 2 + dy:2 + dy + h, 2 + dx:2 + dx + w, :] @ packed[t].float().t()
    return out


def test_conv3x3_packing_is_the_convolution():
    g = torch.Generator().manual_seed(0)
    x = torch.randn((1, 6, 7, 8), generator=g)
    wt = torch.randn((5, 8, 3, 3), generator=g)
    taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    got = _tap_list_conv(x, E.pack_conv3x3(wt).float(), taps).permute(0, 3, 1, 2)
    want = F.conv2d(x.permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), padding=1)
    assert torch.allclose(got, want, atol=1e-4)


def test_convT_phase_decomposition_is_the_transposed_convolution():
    """The 4 phases x 4 taps the kernel runs (tap offsets as in conv_tcgen05.cu snb_conv_create) equal
    nn.ConvTranspose2d(k=4, s=2, p=1)."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn((2, 5, 6, 8), generator=g)
    wt = torch.randn((8, 4, 4, 4), generator=g)
    packed = E.pack_convT4x4(wt).float()
    dlist = [[0, -1], [1, 0]]
    out = torch.zeros((2, 10, 12, 4))
    for py in range(2):
        for px in range(2):
            ph = py * 2 + px
            taps = [(dlist[py][ty], dlist[px][tx]) for ty in range(2) for tx in range(2)]
            out[:, py::2, px::2, :] = _tap_list_conv(x, packed[ph * 4:ph * 4 + 4], taps)
    want = F.conv_transpose2d(x.permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), stride=2, padding=1)
    assert torch.allclose(out.permute(0, 3, 1, 2), want, atol=1e-4)


def test_first_layer_packing_matches_patch_rows():
    g = torch.Generator().manual_seed(2)
    x = torch.randn((1, 3, 5, 6), generator=g)
    wt = torch.randn((4, 3, 3, 3), generator=g)
    cols = F.unfold(x, 3, padding=1).reshape(1, 3, 9, 5, 6).permute(0, 3, 4, 2, 1).reshape(1, 5, 6, 27)
    rows = torch.zeros((1, 5, 6, 32))
    rows[..., :27] = cols
    got = rows @ E.pack_first_conv3x3(wt)[0].float().t()
    want = F.conv2d(x, wt.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(got, want, atol=1e-4)


@pytest.mark.parametrize("arch,cls,n_keys,n_params", [("unet16", UNet16, 76, 32202337), ("unet11", UNet11, 56, 25364513)])
def test_model_mirrors_accept_reference_state_dict(arch, cls, n_keys, n_params):
    m = cls()
    sd = synth.vgg_unet_state_dict(arch, seed=1)
    assert len(m.state_dict()) == n_keys and sorted(m.state_dict()) == sorted(sd)
    m.load_state_dict(sd, strict=True)
    assert sum(p.numel() for p in m.parameters()) == n_params and m.num_classes == 1
    # aliased encoder entries share storage, as in the reference (conv1.0 is encoder.0)
    assert m.state_dict()["conv1.0.weight"].data_ptr() == m.state_dict()["encoder.0.weight"].data_ptr()


def test_d4_helpers_roundtrip_on_cpu_tensors(golden_dir):
    g = np.load(os.path.join(golden_dir, "tta.npz"))
    views = aug.tta_d4_aug([torch.from_numpy(t) for t in g["tiles"]])
    assert np.array_equal(torch.stack(views).numpy(), g["views"])
    deaug = aug.tta_d4_deaug([torch.from_numpy(p) for p in g["preds"]])
    assert np.array_equal(torch.stack(deaug).numpy(), g["deaug"])


def test_shard_range_partitions_everything():
    for n, world in [(180, 8), (169, 4), (5, 8), (0, 2), (7, 1)]:
        spans = [sdist.shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
    assert [e - b for b, e in (sdist.shard_range(180, r, 8) for r in range(8))] == [23, 23, 23, 23, 22, 22, 22, 22]
    with pytest.raises(ValueError):
        sdist.shard_range(4, 2, 2)


def _worker(rank, world, port, n_images, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sdist.init_from_env(backend="gloo")
    b, e = sdist.shard_range(n_images, rank, world)
    counts = torch.zeros(4, dtype=torch.int64)
    masks = []
    for i in range(b, e):
        logits, targets = synth.logits_targets(50 + i, (1, 1, 24, 20))
        p = torch.sigmoid(logits)
        counts += no.confusion_counts(p, targets)
        masks.append(((p > 0.5) * 255).to(torch.uint8).reshape(24, 20))
    local = torch.stack(masks) if masks else torch.zeros((0, 24, 20), dtype=torch.uint8)
    total = sdist.allreduce_counts(counts)
    gathered = sdist.gather_masks(local, n_images)
    if rank == 0:
        torch.save({"counts": total, "masks": gathered}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_counts_and_masks_equal_single_process(tmp_path):
    import torch.multiprocessing as mp

    n_images = 5                                       # uneven split: 3 + 2
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, n_images, out), nprocs=2, join=True)
    got = torch.load(out)
    counts = torch.zeros(4, dtype=torch.int64)
    masks = []
    for i in range(n_images):
        logits, targets = synth.logits_targets(50 + i, (1, 1, 24, 20))
        p = torch.sigmoid(logits)
        counts += no.confusion_counts(p, targets)
        masks.append(((p > 0.5) * 255).to(torch.uint8).reshape(24, 20))
    assert got["counts"].tolist() == counts.tolist()
    assert torch.equal(got["masks"], torch.stack(masks))


def test_zf_unet_mirror_accepts_reference_state_dict():
    m = ZF_UNET()
    sd = synth.zf_unet_state_dict(seed=4)
    assert len(m.state_dict()) == 156 and sorted(m.state_dict()) == sorted(sd)
    m.load_state_dict(sd, strict=True)
    assert sum(p.numel() for p in m.parameters()) == 31454721 and m.num_classes == 1


def test_bn_folding_is_exact_algebra():
    g = torch.Generator().manual_seed(3)
    x = torch.randn((2, 8, 9, 7), generator=g)
    w, b = torch.randn((5, 8, 3, 3), generator=g), torch.randn(5, generator=g)
    gamma, beta = torch.rand(5, generator=g) + 0.5, torch.randn(5, generator=g)
    mean, var = torch.randn(5, generator=g) * 0.1, torch.rand(5, generator=g) + 0.5
    wf, bf = E.fold_bn(w, b, (gamma, beta, mean, var, 1e-5))
    want = F.batch_norm(F.conv2d(x, w, b, padding=1), mean, var, gamma, beta, training=False, eps=1e-5)
    assert torch.allclose(F.conv2d(x, wf, bf, padding=1), want, atol=1e-4)


def test_fcdensenet67_mirror_accepts_reference_state_dict():
    m = FCDenseNet67(n_classes=1)
    sd = synth.fcdensenet_state_dict(seed=5)
    assert len(m.state_dict()) == 434 and sorted(m.state_dict()) == sorted(sd)
    m.load_state_dict(sd, strict=True)
    assert sum(p.numel() for p in m.parameters()) == 3460353 and m.num_classes == 1


def test_convT3x3_phase_decomposition_is_the_cropped_transposed_convolution():
    """Tap slots of SNB_CONVT_3X3_S2 (conv_tcgen05.cu) equal ConvTranspose2d(k=3, s=2, p=0) cropped to [0, 2h) x [0, 2w)."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn((2, 5, 6, 8), generator=g)
    wt = torch.randn((8, 4, 3, 3), generator=g)
    packed = E.pack_convT3x3(wt, 8, 4).float()
    dlist = [[0, -1], [0, 0]]
    out = torch.zeros((2, 10, 12, 4))
    for py in range(2):
        for px in range(2):
            ph = py * 2 + px
            taps = [(dlist[py][ty], dlist[px][tx]) for ty in range(2) for tx in range(2)]
            out[:, py::2, px::2, :] = _tap_list_conv(x, packed[ph * 4:ph * 4 + 4], taps)
    want = F.conv_transpose2d(x.permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), stride=2)[:, :, :10, :12]
    assert torch.allclose(out.permute(0, 3, 1, 2), want, atol=1e-4)


def test_fma_division_by_invariant_is_correctly_rounded():
    """The merge kernel divides by a loop-invariant norm with y = RN(1/b), q0 = RN(a*y) and two FMA corrections
    (csrc/slicer.cu div_by_invariant).  Emulate that sequence with exact rational arithmetic and compare with the
    correctly rounded quotient numpy computes (lib/tiles.py:159), on weighted sums shaped like the merge's."""
    import random
    import struct
    from fractions import Fraction as Fr

    def fma(a, b, c):
        return float(Fr(a) * Fr(b) + Fr(c))          # Fraction -> float rounds to nearest even

    rnd = random.Random(7)
    for i in range(6000):
        if i % 3:
            k = rnd.randint(1, 4)
            ws = [rnd.uniform(0.006, 3.2) for _ in range(k)]
            vs = [float(np.float32(rnd.random())) for _ in range(k)]
            a = b = 0.0
            for v, w in zip(vs, ws):
                a, b = a + v * w, b + w
        else:                                         # arbitrary significands
            a = struct.unpack("d", struct.pack("Q", (rnd.randint(900, 1100) << 52) | rnd.getrandbits(52)))[0]
            b = struct.unpack("d", struct.pack("Q", (rnd.randint(1000, 1040) << 52) | rnd.getrandbits(52)))[0]
        y = float(1 / Fr(b))
        q = float(Fr(a) * Fr(y))
        for _ in range(2):
            q = fma(fma(-b, q, a), y, q)
        assert q == a / b, (a, b)


def _space_to_depth(x_nhwc):
    """snb_space_to_depth2 on the host: out[n][y][x][(py*2+px)*C + c] = in[n][2y+py][2x+px][c]."""
    n, h, w, c = x_nhwc.shape
    return x_nhwc.reshape(n, h // 2, 2, w // 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, h // 2, w // 2, 4 * c)


def test_stride2_conv3x3_packing_is_the_strided_convolution():
    """resnet34 down-sampling block: conv3x3(stride 2, padding 1) == the 4-tap conv (dy, dx in {-1, 0}) the kernel runs over
    the space-to-depth tensor with pack_conv3x3_s2 weights; the 1x1/s2 shortcut == conv1x1 over the first quarter."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn((2, 8, 10, 6), generator=g)                       # NHWC
    wt = torch.randn((5, 6, 3, 3), generator=g)
    taps = [(ty - 1, tx - 1) for ty in range(2) for tx in range(2)]
    got = _tap_list_conv(_space_to_depth(x), E.pack_conv3x3_s2(wt).float(), taps).permute(0, 3, 1, 2)
    want = F.conv2d(x.permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), stride=2, padding=1)
    assert torch.allclose(got, want, atol=1e-4)
    w1 = torch.randn((5, 6, 1, 1), generator=g)
    got1 = _space_to_depth(x)[..., :6] @ E.pack_conv1x1(w1)[0].float().t()
    assert torch.allclose(got1.permute(0, 3, 1, 2), F.conv2d(x.permute(0, 3, 1, 2), w1.to(torch.bfloat16).float(), stride=2), atol=1e-4)


def test_linknet_head_and_stem_packing():
    """conv k2 p1 (finalconv3) as 4 taps over a grid one larger than the input; the 7x7/s2 stem as a GEMM over im2col rows
    with k = (ky*7+kx)*C + c (snb_stem7x7_rows / pack_stem7x7)."""
    g = torch.Generator().manual_seed(4)
    x = torch.randn((1, 5, 7, 4), generator=g)
    wt = torch.randn((3, 4, 2, 2), generator=g)
    xp = F.pad(x, (0, 0, 0, 1, 0, 1))                                   # the kernel's tile grid is (h+1) x (w+1)
    taps = [(ky - 1, kx - 1) for ky in range(2) for kx in range(2)]
    got = _tap_list_conv(xp, E.pack_conv2x2(wt).float(), taps).permute(0, 3, 1, 2)
    assert torch.allclose(got, F.conv2d(x.permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), padding=1), atol=1e-4)
    img = torch.randn((2, 3, 12, 16), generator=g)
    ws = torch.randn((6, 3, 7, 7), generator=g)
    cols = F.unfold(img, 7, padding=3, stride=2).reshape(2, 3, 49, 6, 8).permute(0, 3, 4, 2, 1).reshape(2, 6, 8, 147)
    rows = torch.zeros((2, 6, 8, 160))
    rows[..., :147] = cols
    got = rows @ E.pack_stem7x7(ws, 160)[0].float().t()
    want = F.conv2d(img, ws.to(torch.bfloat16).float(), stride=2, padding=3).permute(0, 2, 3, 1)
    assert torch.allclose(got, want, atol=1e-3)


def test_scatter_packing_sums_to_the_convolution():
    """FCDenseNet growth-rate layer as one N = 144 GEMM (csrc/conv_scatter.cu): P[p][tap*16+co] = x[p] . W[co][:, tap], then
    out[q] = sum_tap P[q + (dy, dx)][tap] -- restated on the host with pack_conv3x3_scatter."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn((1, 9, 11, 32), generator=g)
    wt = torch.randn((16, 32, 3, 3), generator=g)
    P = (x @ E.pack_conv3x3_scatter(wt).float().t()).reshape(1, 9, 11, 9, 16)        # [n][y][x][tap][co]
    Pp = F.pad(P, (0, 0, 0, 0, 1, 1, 1, 1))
    out = sum(Pp[:, ky:ky + 9, kx:kx + 11, ky * 3 + kx] for ky in range(3) for kx in range(3))
    want = F.conv2d(x.permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(out, want, atol=1e-4)


def test_adjoint_weights_give_the_input_gradient():
    """The training plan computes the input gradient of a stride-1 conv3x3 / conv1x1 with the FORWARD kernel on adjoint
    weights (taps flipped, Cin / Cout swapped): check the transform against torch autograd."""
    g = torch.Generator().manual_seed(6)
    for k in (3, 1):
        x = torch.randn((2, 6, 7, 9), generator=g, requires_grad=True)
        wt = torch.randn((5, 6, k, k), generator=g)
        y = F.conv2d(x, wt, padding=k // 2)
        dy = torch.randn(y.shape, generator=g)
        y.backward(dy)
        adj = wt.flip(2, 3).transpose(0, 1).contiguous()
        assert torch.allclose(F.conv2d(dy, adj, padding=k // 2), x.grad, atol=1e-4)


def test_train_plan_host_logic_dry_run():
    """The LinkNet34 training plan's host side (op lists, gradient-slab bookkeeping, index maps of the pack functions,
    gather tables) built and walked on the CPU against a stub of the native library (tools/dry_run_train_plan.py)."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "dry_run_train_plan.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "'WgradOp': 51" in r.stdout and "'ConvOp': 50" in r.stdout and "unpack segments 51" in r.stdout, r.stdout


def _seam_worker(rank, world, port, out):
    """Tile-sharded mode on gloo: each rank 'computes' its crop range (oracle tiles), receives the seam tiles its row band
    needs, merges its band with the oracle merge restricted to those rows; MaskExchange gathers / all-reduces on CPU."""
    import torch.distributed as dist

    from snb_b200 import dist as sdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    h, w, T, S = 300, 260, 128, 96
    s = to.SlicerOracle((h, w, 1), T, S, weight="pyramid")
    rs = np.random.RandomState(7)
    all_tiles = torch.from_numpy(rs.rand(len(s.crops), T, T, 1).astype(np.float32))     # what the network would produce
    tiles_x = len({c[0] for c in s.crops})
    tiles_y = len(s.crops) // tiles_x
    owned = [sdist.shard_range(len(s.crops), r, world) for r in range(world)]
    bands = [sdist.band_range(h, r, world) for r in range(world)]
    needed = [sdist.tiles_covering_rows(b, e, s.margin_top, T, S, tiles_x, tiles_y) for b, e in bands]
    mine = torch.full_like(all_tiles, float("nan"))
    b, e = owned[rank]
    mine[b:e] = all_tiles[b:e]
    sdist.exchange_seam_tiles(mine, owned, needed, rank, world)
    nb, ne = needed[rank]
    assert torch.equal(mine[nb:ne], all_tiles[nb:ne])                    # every tile my band reads has arrived
    rb, re = bands[rank]
    # tiles outside `needed` never touch rows [rb, re): zero them (NaN would poison the oracle's full-image accumulate)
    safe = torch.where(torch.isnan(mine), torch.zeros(()), mine)
    band = torch.from_numpy(s.merge(list(safe.numpy()), dtype=np.float32)[rb:re])
    mask = ((band > 0.5) * 255).to(torch.uint8)
    ex = sdist.MaskExchange((bands[0][1] - bands[0][0], w, 1), torch.device("cpu"))
    pad = torch.zeros((bands[0][1] - bands[0][0], w, 1), dtype=torch.uint8)
    pad[:re - rb] = mask
    counts = torch.tensor([int((mask > 0).sum()), 0, 0, int((mask == 0).sum())])
    slot = ex.submit(pad, counts)
    ex.wait()
    total, gathered = ex.result(slot)
    if rank == 0:
        full = torch.cat([gathered[r][:bands[r][1] - bands[r][0]] for r in range(world)])
        torch.save({"mask": full, "counts": total.clone(), "band0": band}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_tile_sharded_bands_equal_single_process(tmp_path):
    """SURVEY 8e 'shard by tile': seam-tile exchange + per-rank band merge + mask gather reproduce the single-process
    merge byte for byte (gloo, CPU tensors; the device path runs the same dist.py functions over NCCL)."""
    import torch.multiprocessing as mp

    from snb_b200 import dist as sdist

    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    out = str(tmp_path / "seam.pt")
    mp.spawn(_seam_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    h, w, T, S = 300, 260, 128, 96
    s = to.SlicerOracle((h, w, 1), T, S, weight="pyramid")
    tiles = np.random.RandomState(7).rand(len(s.crops), T, T, 1).astype(np.float32)
    want = s.merge(list(tiles), dtype=np.float32)
    assert np.array_equal(got["band0"].numpy(), want[:150])
    assert np.array_equal(got["mask"].numpy(), ((want > 0.5) * 255).astype(np.uint8))
    assert got["counts"].tolist() == [int((want > 0.5).sum()), 0, 0, int((want <= 0.5).sum())]
    # helper properties
    assert sdist.pick_tile_batch(169) == 85 and sdist.pick_tile_batch(22) == 22 and sdist.pick_tile_batch(85) == 85 and sdist.pick_tile_batch(338) == 85
    assert sdist.range_overlap((0, 10), (7, 20)) == (7, 3) and sdist.range_overlap((0, 5), (7, 9))[1] == 0
    for world in (1, 2, 3, 8):
        rows = [sdist.band_range(5000, r, world) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == 5000 and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
