"""Multi-rank byte-equality check (run under torchrun, NCCL, one process per GPU): the image-sharded job (MaskExchange:
all-reduce of int64 counts + gather of uint8 masks) and the tile-sharded job (TileShardedPredictor: seam-tile send / recv,
band merge, mask all-gather) must reproduce the single-GPU masks BYTE FOR BYTE and the same int64 counts (SURVEY 8e,
BASELINE configs[3]).  Every rank also runs the whole job alone as the reference.  Prints MULTIRANK_OK on rank 0."""
import argparse
import os
import sys

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
ap.add_argument("--images", type=int, default=5)
ap.add_argument("--h", type=int, default=1100)
ap.add_argument("--w", type=int, default=900)
ap.add_argument("--tile", type=int, default=256)
ap.add_argument("--step", type=int, default=192)
args = ap.parse_args()
rank, world, local = sdist.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
m = UNet16()
m.load_state_dict(synth.vgg_unet_state_dict("unet16", seed=0))
m = m.to(dev).eval()
shape = (args.h, args.w, 3)
images = [torch.from_numpy(synth.image_u8(i, args.h, args.w)).to(dev) for i in range(args.images)]
gts = [torch.from_numpy(synth.gt_mask_u8(i, args.h, args.w)).to(dev).reshape(args.h, args.w, 1) for i in range(args.images)]

# ---- single-GPU reference on every rank
solo = sub.TiledPredictor(m, shape, args.tile, args.step, batch_size=8, tta=False, device=dev)
ref_masks, ref_merged, ref_counts = [], [], torch.zeros(4, dtype=torch.int64, device=dev)
for img, gt in zip(images, gts):
    merged, mask = solo.predict_device(img)
    ref_masks.append(mask.clone())
    ref_merged.append(merged.clone())
    ref_counts += metrics.confusion_counts_from_probs(merged, gt)
ok = True

# ---- image-sharded: contiguous image ranges per rank, masks gathered on rank 0, counts all-reduced
b, e = sdist.shard_range(args.images, rank, world)
ex = sdist.MaskExchange((args.h, args.w, 1), dev)
slots, total = [], torch.zeros(4, dtype=torch.int64, device=dev)
per_rank = max(sdist.shard_range(args.images, r, world)[1] - sdist.shard_range(args.images, r, world)[0] for r in range(world))
got = {}
for k in range(per_rank):
    i = b + k
    if i < e:
        merged, mask = solo.predict_device(images[i])
        counts = metrics.confusion_counts_from_probs(merged, gts[i])
    else:                           # ranks with fewer images still take part in the collectives
        mask, counts = torch.zeros_like(ref_masks[0]), torch.zeros(4, dtype=torch.int64, device=dev)
    slot = ex.submit(mask, counts)
    ex.wait()
    c, g = ex.result(slot)
    total += c
    if rank == 0:
        for r in range(world):
            rb, re = sdist.shard_range(args.images, r, world)
            if rb + k < re:
                got[rb + k] = g[r].clone() if world > 1 else g.clone()
if rank == 0:
    ok &= sorted(got) == list(range(args.images)) and all(torch.equal(got[i], ref_masks[i]) for i in range(args.images))
ok &= total.tolist() == ref_counts.tolist()

# ---- tile-sharded: every image split by crop range over all ranks
tsp = sub.TileShardedPredictor(m, shape, args.tile, args.step, tta=False, device=dev, gather_probs=True, overlap=True)
tcounts = torch.zeros(4, dtype=torch.int64, device=dev)
for i, (img, gt) in enumerate(zip(images, gts)):
    merged, mask, counts = tsp.predict_device(img, gt)
    tsp.wait()
    ok &= torch.equal(mask, ref_masks[i]) and torch.equal(merged, ref_merged[i])
    tcounts += counts
ok &= tcounts.tolist() == ref_counts.tolist()
# pipelined use (no wait between images): the last image's result is still right
for img, gt in zip(images, gts):
    merged, mask, counts = tsp.predict_device(img, gt)
tsp.wait()
ok &= torch.equal(mask, ref_masks[-1])

flag = torch.tensor([1 if ok else 0], device=dev)
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIRANK_OK" if int(flag) == 1 else "MULTIRANK_MISMATCH", "world", world, "images", args.images,
          "counts", ref_counts.tolist(), "seam bytes per image and rank", tsp.exchange_bytes, flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
