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
