"""BASELINE configs[3]: AlbuNet (UNet16) tiled inference over 180 synthetic Inria images (5000 x 5000, 512 / 384) sharded over
the GPUs of one box, NCCL all-reduce of the int64 IoU counts, uint8 masks gathered on rank 0 (SURVEY 8d "Config 4", 8e).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_config3.py [--images 180]
        [--shard image|tile] [--verify]

--shard image (default): contiguous image ranges per rank, every rank merges its own images (byte-identical to one GPU).
--shard tile: every image split by crop range over all ranks (seam tiles by send / recv, band merge, mask all-gather).
--verify: rank 0 also runs every image alone and checks that all gathered masks are BYTE-identical and the counts equal.
Images are RandomState(i).randint(0, 256, (5000, 5000, 3)), ground truth RandomState(1000 + i).rand(5000, 5000) > 0.5.
Prints one JSON line on rank 0.  The gathered masks stay on the device (180 x 25 MB = 4.5 GB of rank 0's HBM).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snb_b200  # noqa: E402,F401
from snb_b200 import dist as sdist  # noqa: E402
from snb_b200 import inria_submit as sub  # noqa: E402
from snb_b200 import synth  # noqa: E402
from snb_b200.lib import metrics  # noqa: E402
from snb_b200.lib.models import UNet16  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=180)
ap.add_argument("--size", type=int, default=5000)
ap.add_argument("--tile", type=int, default=512)
ap.add_argument("--step", type=int, default=384)
ap.add_argument("--shard", default="image", choices=["image", "tile"])
ap.add_argument("--verify", action="store_true")
args = ap.parse_args()
rank, world, local = sdist.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
H = W = args.size
shape = (H, W, 3)


def image(i):
    return torch.from_numpy(np.random.RandomState(i).randint(0, 256, shape).astype(np.uint8))


def truth(i):
    return torch.from_numpy((np.random.RandomState(1000 + i).rand(H, W) > 0.5).astype(np.uint8)).reshape(H, W, 1)


model = UNet16()
model.load_state_dict(synth.vgg_unet_state_dict("unet16", seed=0))
model = model.to(dev).eval()
total = torch.zeros(4, dtype=torch.int64, device=dev)
all_masks = torch.empty((args.images, H, W, 1), dtype=torch.uint8, device=dev) if rank == 0 else None
t_gen = 0.0
torch.cuda.synchronize()
t0 = time.perf_counter()
if args.shard == "image":
    pred = sub.TiledPredictor(model, shape, args.tile, args.step, batch_size=13, tta=False, device=dev)
    ex = sdist.MaskExchange((H, W, 1), dev)
    b, e = sdist.shard_range(args.images, rank, world)
    ranges = [sdist.shard_range(args.images, r, world) for r in range(world)]
    steps = max(re - rb for rb, re in ranges)
    for k in range(steps):
        i = b + k
        if i < e:
            tg = time.perf_counter()
            img, gt = image(i).to(dev), truth(i).to(dev)
            t_gen += time.perf_counter() - tg
            merged, mask = pred.predict_device(img)
            counts = metrics.confusion_counts_from_probs(merged, gt)
        else:                                   # a rank with one image less still takes part in the collectives
            mask, counts = torch.zeros((H, W, 1), dtype=torch.uint8, device=dev), torch.zeros(4, dtype=torch.int64, device=dev)
        slot = ex.submit(mask, counts)
        ex.wait()
        c, g = ex.result(slot)
        total += c
        if rank == 0:
            for r, (rb, re) in enumerate(ranges):
                if rb + k < re:
                    all_masks[rb + k].copy_(g[r] if world > 1 else g)
else:
    pred = sub.TileShardedPredictor(model, shape, args.tile, args.step, tta=False, device=dev, overlap=False)
    for i in range(args.images):
        tg = time.perf_counter()
        img, gt = image(i).to(dev), truth(i).to(dev)          # every rank needs the image; its band of the truth would do
        t_gen += time.perf_counter() - tg
        merged, mask, counts = pred.predict_device(img, gt)
        total += counts
        if rank == 0:
            all_masks[i].copy_(mask)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
wall = time.perf_counter() - t0

ok = None
if args.verify and rank == 0:
    solo = sub.TiledPredictor(model, shape, args.tile, args.step, batch_size=13, tta=False, device=dev)
    ref = torch.zeros(4, dtype=torch.int64, device=dev)
    ok = True
    for i in range(args.images):
        merged, mask = solo.predict_device(image(i).to(dev))
        ref += metrics.confusion_counts_from_probs(merged, truth(i).to(dev))
        ok = ok and bool(torch.equal(mask, all_masks[i]))
    ok = ok and ref.tolist() == total.tolist()
if rank == 0:
    tp, fp, fn, tn = total.tolist()
    print(json.dumps({"config": "configs[3]: %d images %dx%d, tile %d / step %d, shard by %s over %d GPUs" % (
        args.images, H, W, args.tile, args.step, args.shard, world), "counts_tp_fp_fn_tn": [tp, fp, fn, tn],
        "iou": tp / max(1, tp + fp + fn), "pixels": tp + fp + fn + tn, "wall_s": wall, "host_image_synthesis_s_rank0": t_gen,
        "mpx_per_s_wall": args.images * H * W / 1e6 / wall, "masks_and_counts_identical_to_single_gpu": ok}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
