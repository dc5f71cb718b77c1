"""One-process-per-GPU sharding of the tiled-inference job: images are independent, so each rank owns a
contiguous range of images and merges them with the single-GPU accumulation order (byte-identical masks); the only
exchanges are an all-reduce of the int64 confusion counts and a gather of the uint8 masks.
Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style init (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*); returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_range(n_items, rank, world):
    """Contiguous [begin, end) of `n_items` for `rank`: the first n % world ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def band_range(n_rows, rank, world):
    """Contiguous band [begin, end) of image rows owned by `rank` in the tile-sharded mode (equal split)."""
    return shard_range(n_rows, rank, world)


def tiles_covering_rows(row_begin, row_end, margin_top, tile, step, tiles_x, tiles_y):
    """Crop range [begin, end) (crop order: y outer, x inner, lib/tiles.py:94-96) of all crops whose rows overlap the image
    rows [row_begin, row_end): crop row ky covers image rows [ky * step - margin_top, ky * step - margin_top + tile)."""
    if row_end <= row_begin:
        return 0, 0
    ky_lo = max(0, -(-(row_begin + margin_top - tile + 1) // step))          # first ky with ky*step - mt + tile > row_begin
    ky_hi = min(tiles_y - 1, (row_end - 1 + margin_top) // step)               # last ky with ky*step - mt <= row_end - 1
    return ky_lo * tiles_x, (ky_hi + 1) * tiles_x


def range_overlap(a, b):
    """(begin, count) of the intersection of two [begin, end) ranges"""
    lo, hi = max(a[0], b[0]), min(a[1], b[1])
    return lo, max(0, hi - lo)


def exchange_seam_tiles(tiles, owned, needed, rank, world):
    """Tile-sharded mode: `tiles` is the [n_tiles, ...] store of every rank, owned[r] the crop range rank r computed,
    needed[r] the crop range rank r's row band reads.  What I own and a peer needs goes out, what I need and a peer owns
    comes in -- one grouped send / recv (NCCL: a single kernel), straight between the tile stores."""
    ops = []
    for peer in range(world):
        if peer == rank:
            continue
        lo, cnt = range_overlap(owned[rank], needed[peer])
        if cnt:
            ops.append(dist.P2POp(dist.isend, tiles[lo:lo + cnt], peer))
        lo, cnt = range_overlap(owned[peer], needed[rank])
        if cnt:
            ops.append(dist.P2POp(dist.irecv, tiles[lo:lo + cnt], peer))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def pick_tile_batch(n_tiles, preferred=85, lo=24, hi=96):
    """Tiles per network launch for a shard of n_tiles: the plan always runs whole batches, so the batch that wastes the
    fewest padded tile slots wins (ties: closest to the preferred size).  Large batches (measured on UNet16 at 512 x 512:
    85 tiles per launch +3 % over 13): the low-resolution layers then fill many waves of 148 CTAs instead of 5.6."""
    if n_tiles <= hi:
        return max(1, n_tiles)
    slack = max(1, n_tiles // 100)        # padded slots that are as good as none
    best = None
    for b in range(lo, hi + 1):
        waste = -(-n_tiles // b) * b - n_tiles
        key = (max(0, waste - slack), abs(b - preferred), waste)
        if best is None or key < best[0]:
            best = (key, b)
    return best[1]


class MaskExchange:
    """The only exchange of the image-sharded job, off the compute stream: all-reduce of the int64 [tp, fp, fn,
    tn] counts and gather of the uint8 masks to `dst`, on a side stream with preallocated double buffers, so the NCCL
    traffic of image i overlaps the convolutions of image i+1.  `submit(mask, counts)` is called on the compute stream
    after an image is done; `wait()` makes the current stream wait for everything submitted."""

    def __init__(self, mask_shape, device, dst=0):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.dst, self.device = dst, device
        self.stream = torch.cuda.Stream(device=device) if device.type == 'cuda' else None
        self.masks = [torch.empty(mask_shape, dtype=torch.uint8, device=device) for _ in range(2)]
        self.counts = [torch.zeros(4, dtype=torch.int64, device=device) for _ in range(2)]
        self.gathered = [None, None]
        if self.rank == dst and self.world > 1:
            self.gathered = [torch.empty((self.world,) + tuple(mask_shape), dtype=torch.uint8, device=device) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)] if self.stream else None
        self.done = [torch.cuda.Event() for _ in range(2)] if self.stream else None
        self.step = 0

    def submit(self, mask, counts):
        slot = self.step & 1
        self.step += 1
        if self.stream is None:                          # CPU / gloo: same calls, in line
            self.masks[slot].copy_(mask.reshape(self.masks[slot].shape))
            self.counts[slot].copy_(counts)
            self._exchange(slot)
            return slot
        cur = torch.cuda.current_stream(self.device)
        if self.step > 2:
            cur.wait_event(self.done[slot])              # the exchange that used this slot two images ago has finished
        self.masks[slot].copy_(mask.reshape(self.masks[slot].shape), non_blocking=True)
        self.counts[slot].copy_(counts, non_blocking=True)
        self.ready[slot].record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.ready[slot])
            self._exchange(slot)
            self.done[slot].record(self.stream)
        return slot

    def _exchange(self, slot):
        if self.world > 1:
            dist.all_reduce(self.counts[slot], op=dist.ReduceOp.SUM)
            glist = list(self.gathered[slot].unbind(0)) if self.rank == self.dst else None
            dist.gather(self.masks[slot], glist, dst=self.dst)

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def result(self, slot):
        """(all-reduced counts, [world, ...] masks on dst / the local mask elsewhere) of a submitted slot, after wait()."""
        g = self.gathered[slot]
        return self.counts[slot], (g if g is not None else self.masks[slot])


def allreduce_counts(counts):
    """Sum of the int64 [tp, fp, fn, tn] vectors over all ranks (exact: integer addition is order independent)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def gather_masks(local_masks, n_total, dst=0):
    """Gather per-rank uint8 mask stacks [n_local, H, W] into [n_total, H, W] on `dst` (image order preserved).

    Ranks may own different image counts (shard_range), so every rank pads to the largest shard."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_masks
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    biggest = max(e - b for b, e in sizes)
    pad = torch.zeros((biggest,) + tuple(local_masks.shape[1:]), dtype=local_masks.dtype, device=local_masks.device)
    pad[:local_masks.shape[0]] = local_masks
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bufs[r][:e - b] for r, (b, e) in enumerate(sizes)], dim=0)
