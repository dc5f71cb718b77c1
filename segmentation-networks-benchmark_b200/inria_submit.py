"""Tiled full-resolution inference, the path reference inria_submit.py:237-306 drives, kept on the device.

`predict_tiled(image, model, test_transform, patch_size, batch_size)` keeps the reference signature; underneath,
`TiledPredictor` holds the whole per-image pipeline on one GPU:

    uint8 HWC image (HBM) --snb_split_norm_u8--> first-layer operand rows of a batch of tiles
        --VGGUNetPlan.run (tcgen05 convs, sigmoid fused into the last epilogue)--> float32 probabilities
        --snb_merge (D4 de-augmentation, float64 pyramid-weighted overlap-add, threshold)--> float32 map + uint8 mask

so an image crosses PCIe once in (75 MB) and once out (25 MB) instead of once per batch in each direction
(inria_submit.py:249,251).  The reference hard-codes tile_step = patch_size // 2 and no way to switch TTA off
(inria_submit.py:240,243); both are parameters here, defaulting to the reference behaviour.
"""
import numpy as np
import torch

from . import _native as N
from . import engine
from .lib.augmentations import NormalizeImage, find_normalize
from .lib.tiles import ImageSlicer

INRIA_MEAN = [0.40273115, 0.45046371, 0.42960134]   # reference lib/datasets/Inria.py:34-35
INRIA_STD = [3.15086464, 3.29831641, 3.63201004]


class TiledPredictor:
    def __init__(self, model, image_shape, patch_size, tile_step=None, batch_size=1, weight='pyramid', tta=True,
                 normalize=None, device=None, use_graph=True, tile_range=None, merge=True):
        N.require_cuda()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.image_shape = tuple(image_shape)
        self.channels = image_shape[2] if len(image_shape) == 3 else 1
        self.patch_size = patch_size
        self.tile_step = patch_size // 2 if tile_step is None else tile_step
        self.views = 8 if tta else 1
        self.slicer = ImageSlicer(image_shape, patch_size, self.tile_step, weight=weight)
        self.n_tiles = len(self.slicer.crops)
        # tile_range = (begin, end): this predictor only runs those crops (tile-sharded multi-GPU mode); with merge=False
        # the caller exchanges self.probs and calls merge_probs() itself
        self.tile_begin, self.tile_end = (0, self.n_tiles) if tile_range is None else tile_range
        if not (0 <= self.tile_begin <= self.tile_end <= self.n_tiles):
            raise ValueError("tile_range outside the %d crops" % self.n_tiles)
        self.do_merge = merge
        self.batch = max(1, min(int(batch_size), max(1, self.tile_end - self.tile_begin)))
        norm = normalize if normalize is not None else NormalizeImage(mean=INRIA_MEAN, std=INRIA_STD)
        with torch.cuda.device(self.device):
            self.lut = torch.from_numpy(norm.lut()).to(self.device)
            self.plan = model.plan(self.batch, patch_size, patch_size, sigmoid=True)
            self.weight = self.slicer.weight_on_device(self.device)
            T = patch_size
            # (padded to whole network launches: the last layer writes its batch straight into this buffer)
            n_alloc = max(self.n_tiles, self.tile_begin + -(-(self.tile_end - self.tile_begin) // self.batch) * self.batch)
            self._probs_store = torch.empty((n_alloc, self.views, T, T, 1), dtype=torch.float32, device=self.device)
            self.probs = self._probs_store[:self.n_tiles]
            h, w = image_shape[0], image_shape[1]
            self.merged = torch.empty((h, w, 1), dtype=torch.float32, device=self.device)
            self.mask = torch.empty((h, w, 1), dtype=torch.uint8, device=self.device)
            self.image = torch.empty((h, w, self.channels), dtype=torch.uint8, device=self.device)
        self.use_graph = use_graph
        self._graph = None
        n_local = self.tile_end - self.tile_begin
        n_chunks = (n_local + self.batch - 1) // self.batch
        # kernels enqueued per image: per chunk and view one split + the plan, then one merge
        self.launches_per_image = n_chunks * self.views * (1 + self.plan.launches) + (1 if merge else 0)
        self.flops_per_image = self.plan.flops / self.batch * n_local * self.views

    def predict_device(self, d_image):
        """uint8 [H][W][C] CUDA tensor -> (float32 [H][W][1] merged probabilities, uint8 [H][W][1] mask); async.

        The ~430 kernel launches of one image are captured once into a CUDA graph and replayed (the launch-bound
        inner loop of inria_submit.py:248-253 becomes one graph launch); outputs live in self.merged / self.mask."""
        if d_image.dtype != torch.uint8 or tuple(d_image.shape[:2]) != self.image_shape[:2]:
            raise ValueError("expected a uint8 image of shape %s" % (self.image_shape,))
        self.image.copy_(d_image.reshape(self.image.shape), non_blocking=True)   # H2D when the source is pinned host memory
        if not self.use_graph:
            return self._enqueue(self.image)
        if self._graph is None:
            self._enqueue(self.image)                     # warm-up outside capture (lazy module loading etc.)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue(self.image)
            self._graph = g
        self._graph.replay()
        return self.merged, self.mask

    def _enqueue(self, d_image):
        lib, st = N.lib(), N.stream_ptr()
        # Without TTA a batch of probability tiles is contiguous in self.probs: the fused head of the last layer writes it
        # there directly (no copy kernel per batch).  A partial last batch may only do so when the tile slots it spills into
        # are padding, i.e. this predictor owns the tail of the crop list.
        head = engine.head_op(self.plan) if (self.views == 1 and hasattr(self, "_probs_store")) else None
        for begin in range(self.tile_begin, self.tile_end, self.batch):
            count = min(self.batch, self.tile_end - begin)
            direct = head is not None and (count == self.batch or self.tile_end == self.n_tiles)
            for v in range(self.views):
                x_nchw = getattr(self.plan, "x_nchw", None)
                if x_nchw is not None:      # plans that take normalised float NCHW tiles (LinkNet34's 7x7 stem)
                    layout, target = N.LAYOUT_NCHW_F32, x_nchw.data_ptr()
                else:                       # packed 3-channel bf16 tile (first-layer operand rows built on chip), or PATCH32
                    layout, target = self.plan.input_layout()
                N.check(lib.snb_split_norm_u8(self.slicer.handle, N.ptr(d_image), self.channels, N.ptr(self.lut), v,
                                              layout, N.c_vp(target), begin, count, st))
                if direct:
                    head.set_head_out(self._probs_store[begin].data_ptr())
                    self.plan.run()
                else:
                    out = self.plan.run()
                    self.probs[begin:begin + count, v, :, :, 0].copy_(out[:count])
        if head is not None:
            head.set_head_out(self.plan.out.data_ptr())      # the plan is shared (model.plan cache): leave it as it was
        if self.do_merge:
            self.merge_probs()
        return self.merged, self.mask

    def merge_probs(self):
        """Weighted overlap-add of self.probs (all crops) into self.merged / self.mask on the current stream."""
        N.check(N.lib().snb_merge(self.slicer.handle, N.ptr(self.probs), N.DT_F32, 1, self.views, N.ptr(self.weight),
                                  N.ptr(self.merged), N.DT_F32, N.ptr(self.mask), 0.5, N.stream_ptr()))
        return self.merged, self.mask

    def __call__(self, image):
        """numpy uint8 H x W x C -> numpy float32 H x W x 1 (what reference predict_tiled returns)."""
        merged, _ = self.predict_device(torch.from_numpy(np.ascontiguousarray(image)))
        return merged.cpu().numpy()


class TileShardedPredictor:
    """One image across all ranks (BASELINE configs[3], "sharded by tile"): rank r runs the contiguous crop range
    dist.shard_range(n_tiles, r, world) and OWNS a band of image rows.  Only the seam tiles cross NVLink: a rank receives
    the probability tiles of the crop rows that overlap its band and that a neighbour computed (NCCL grouped send / recv
    straight between the ranks' tile buffers, ~1/8 of what an all-gather of all tiles would move), merges its band with
    snb_merge_rows in crop order -- never partial float accumulators, so the bytes equal the single-GPU merge -- and the
    uint8 mask bands are all-gathered (25 MB per image in total).  Exchange, band merge and gather run on a side stream
    with double-buffered tiles, so they overlap the next image's convolutions.  Strong scaling of single-image latency.

    predict_device(d_image, gt=None) -> (merged, mask[, counts]): `mask` is the full uint8 mask on every rank; `merged`
    holds this rank's band of the float32 map (all of it with gather_probs=True); counts = all-reduced int64 [tp, fp, fn,
    tn] against `gt` (uint8 [H, W, 1] on every rank) when given."""

    def __init__(self, model, image_shape, patch_size, tile_step=None, batch_size=None, weight='pyramid', tta=True,
                 normalize=None, device=None, use_graph=True, gather_probs=False, overlap=False):
        import torch.distributed as dist

        from . import dist as sdist

        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        probe = ImageSlicer(image_shape, patch_size, patch_size // 2 if tile_step is None else tile_step, weight=weight)
        n_tiles = len(probe.crops)
        self.ranges = [sdist.shard_range(n_tiles, r, self.world) for r in range(self.world)]
        b, e = self.ranges[self.rank]
        batch = sdist.pick_tile_batch(e - b) if batch_size is None else batch_size
        self.local = TiledPredictor(model, image_shape, patch_size, tile_step, batch, weight, tta, normalize, device, use_graph,
                                    tile_range=(b, e), merge=False)
        p = self.local
        h = image_shape[0]
        tiles_x = len({x for x, _, _, _ in probe.crops})
        tiles_y = n_tiles // tiles_x
        self.bands = [sdist.band_range(h, r, self.world) for r in range(self.world)]
        self.needs = [sdist.tiles_covering_rows(rb, re, probe.margin_top, patch_size, p.tile_step, tiles_x, tiles_y)
                      for rb, re in self.bands]
        self.max_band = max(re - rb for rb, re in self.bands)
        self.gather_probs = gather_probs
        # overlap=True: predict_device returns while the exchange / merge / gather of this image still run on the side
        # stream (call wait() before reading the results); False: results are ordered on the caller's stream
        self.overlap = overlap
        self.side = torch.cuda.Stream(device=p.device)
        # double-buffered tile store for the side stream (the predictor's own buffer is baked into its CUDA graph)
        self.stage = [torch.empty_like(p.probs) for _ in range(2)]
        w = image_shape[1]
        self.band_mask = torch.zeros((self.max_band, w, 1), dtype=torch.uint8, device=p.device)
        self.all_masks = torch.empty((self.world, self.max_band, w, 1), dtype=torch.uint8, device=p.device)
        self.band_f32 = torch.zeros((self.max_band, w, 1), dtype=torch.float32, device=p.device) if gather_probs else None
        self.all_f32 = torch.empty((self.world, self.max_band, w, 1), dtype=torch.float32, device=p.device) if gather_probs else None
        self.counts = torch.zeros(4, dtype=torch.int64, device=p.device)
        self.computed = [torch.cuda.Event() for _ in range(2)]
        self.merged_ev = [torch.cuda.Event() for _ in range(2)]
        self.step = 0
        self.exchange_bytes = sum(sdist.range_overlap(self.ranges[s], self.needs[self.rank])[1] for s in range(self.world)
                                  if s != self.rank) * p.probs[0].numel() * 4

    def predict_device(self, d_image, gt=None):
        import torch.distributed as dist

        from .lib import metrics

        p = self.local
        slot = self.step & 1
        self.step += 1
        cur = torch.cuda.current_stream(p.device)
        p.predict_device(d_image)                          # split + network for my crop range -> p.probs[b:e]
        if self.step > 2:
            cur.wait_event(self.merged_ev[slot])          # the merge that read this staging buffer two images ago is done
        b, e = self.ranges[self.rank]
        stage = self.stage[slot]
        stage[b:e].copy_(p.probs[b:e], non_blocking=True)
        self.computed[slot].record(cur)
        rb, re = self.bands[self.rank]
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.computed[slot])
            if self.world > 1:
                from . import dist as sdist
                sdist.exchange_seam_tiles(stage, self.ranges, self.needs, self.rank, self.world)
            N.check(N.lib().snb_merge_rows(p.slicer.handle, N.ptr(stage), N.DT_F32, 1, p.views, N.ptr(p.weight),
                                           N.ptr(p.merged), N.DT_F32, N.ptr(p.mask), 0.5, rb, re - rb, N.stream_ptr()))
            self.merged_ev[slot].record(self.side)
            if gt is not None:
                metrics.confusion_counts_from_probs(p.merged[rb:re], gt[rb:re], out=self.counts)
            if self.world > 1:
                self.band_mask[:re - rb].copy_(p.mask[rb:re])
                dist.all_gather_into_tensor(self.all_masks, self.band_mask)
                for r, (b0, b1) in enumerate(self.bands):
                    if r != self.rank:
                        p.mask[b0:b1].copy_(self.all_masks[r, :b1 - b0])
                if self.gather_probs:
                    self.band_f32[:re - rb].copy_(p.merged[rb:re])
                    dist.all_gather_into_tensor(self.all_f32, self.band_f32)
                    for r, (b0, b1) in enumerate(self.bands):
                        if r != self.rank:
                            p.merged[b0:b1].copy_(self.all_f32[r, :b1 - b0])
                if gt is not None:
                    dist.all_reduce(self.counts, op=dist.ReduceOp.SUM)
        if not self.overlap:
            self.wait()
        return (p.merged, p.mask) if gt is None else (p.merged, p.mask, self.counts)

    def wait(self):
        """make the current stream wait for the exchange / merge / gather of everything submitted"""
        torch.cuda.current_stream(self.local.device).wait_stream(self.side)


class StreamingPredictor:
    """Host-to-host pipeline over a sequence of images: while image i runs on the compute stream, image
    i+1 is uploaded and mask i-1 is downloaded on a copy stream from / to pinned host buffers, so PCIe time
    (75 MB in + 25 MB out per 5000x5000 image) hides behind the ~50 ms of compute.  `submit(image)` returns the
    uint8 mask of the PREVIOUS image (None for the first call); `flush()` returns the last one."""

    def __init__(self, predictor):
        self.p = predictor
        dev = predictor.device
        shape = predictor.image.shape
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.stage = [torch.empty(shape, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.mask_dev = [torch.empty_like(predictor.mask) for _ in range(2)]
        self.mask_host = [torch.empty(predictor.mask.shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.uploaded = [torch.cuda.Event() for _ in range(2)]
        self.computed = [torch.cuda.Event() for _ in range(2)]
        self.downloaded = [torch.cuda.Event() for _ in range(2)]
        self.count = 0
        self.used = [False, False]
        self._prefetched = False

    def _upload(self, slot, host_image):
        with torch.cuda.stream(self.copy_stream):
            if self.used[slot]:
                self.copy_stream.wait_event(self.computed[slot])   # the previous image in this slot has been consumed
            self.stage[slot].copy_(host_image.reshape(self.stage[slot].shape), non_blocking=True)
            self.uploaded[slot].record(self.copy_stream)
        self.used[slot] = True

    def submit(self, host_image, next_host_image=None):
        """host_image: pinned uint8 CPU tensor.  Pass next_host_image to start its upload under this image's compute."""
        i = self.count
        slot = i & 1
        cur = torch.cuda.current_stream(self.p.device)
        if not self._prefetched:
            self._upload(slot, host_image)
        cur.wait_event(self.uploaded[slot])
        if next_host_image is not None:
            self._upload(slot ^ 1, next_host_image)
            self._prefetched = True
        else:
            self._prefetched = False
        _, mask = self.p.predict_device(self.stage[slot])
        if i >= 2:
            cur.wait_event(self.downloaded[slot])          # the D2H that read mask_dev[slot] two images ago has finished
        self.mask_dev[slot].copy_(mask, non_blocking=True)   # 25 MB device copy frees the predictor's buffer
        self.computed[slot].record(cur)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.computed[slot])
            self.mask_host[slot].copy_(self.mask_dev[slot], non_blocking=True)
            self.downloaded[slot].record(self.copy_stream)
        self.count += 1
        if i == 0:
            return None
        prev = slot ^ 1
        self.downloaded[prev].synchronize()
        return self.mask_host[prev]

    def flush(self):
        if self.count == 0:
            return None
        last = (self.count - 1) & 1
        self.downloaded[last].synchronize()
        return self.mask_host[last]


class FileSubmitter:
    """The loop of reference inria_submit.main (inria_submit.py:291-306) as a three-stage host pipeline around the device
    pipeline: a pool of DECODE threads (`cv2.imread`, releases the GIL) reads images ahead into pinned
    buffers, the StreamingPredictor overlaps H2D / compute / D2H of consecutive images, and a pool of WRITE threads encodes
    the uint8 masks (`cv2.imwrite`) while the GPU is already on the next images.  With the network at ~50 ms per
    5000 x 5000 image, TIFF decode (~100-300 ms per image on one core) is what bounds a directory run: `decoders` threads
    hide it.  `read` / `write` are injectable (tests, other formats).  All images must share one shape."""

    def __init__(self, predictor, decoders=4, writers=2, prefetch=None, read=None, write=None):
        from concurrent.futures import ThreadPoolExecutor

        from .lib.common import read_rgb

        self.predictor = predictor
        self.streamer = StreamingPredictor(predictor)
        self.read = read_rgb if read is None else read
        self.write = write if write is not None else self._imwrite
        self.decode_pool = ThreadPoolExecutor(max_workers=max(1, decoders), thread_name_prefix="snb-decode")
        self.write_pool = ThreadPoolExecutor(max_workers=max(1, writers), thread_name_prefix="snb-write")
        self.prefetch = max(2, decoders * 2 if prefetch is None else prefetch)
        shape = tuple(predictor.image.shape)
        self._free = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(self.prefetch + 2)]

    @staticmethod
    def _imwrite(path, mask):
        import cv2

        if not cv2.imwrite(path, mask):
            raise IOError("cv2.imwrite failed for %s" % path)

    def _decode(self, path, buf):
        image = self.read(path)
        if image is None:
            raise FileNotFoundError(path)
        if image.ndim == 2:
            image = image[..., None]
        if tuple(image.shape) != tuple(buf.shape):
            raise ValueError("%s has shape %s, the predictor was built for %s" % (path, image.shape, tuple(buf.shape)))
        buf.numpy()[...] = image                      # into pinned memory: the H2D copy is then asynchronous
        return buf

    def run(self, image_paths, out_dir, suffix='.tif'):
        """Predict every file, write `<basename><suffix>` masks into out_dir; returns the list of output paths."""
        import os
        from collections import deque

        os.makedirs(out_dir, exist_ok=True)
        paths = list(image_paths)
        outs = [os.path.join(out_dir, os.path.splitext(os.path.basename(p))[0] + suffix) for p in paths]
        pending, writes, in_gpu = deque(), [], deque()
        nxt = 0

        def feed():
            nonlocal nxt
            while nxt < len(paths) and len(pending) < self.prefetch and self._free:
                buf = self._free.pop()
                pending.append((nxt, buf, self.decode_pool.submit(self._decode, paths[nxt], buf)))
                nxt += 1

        def retire(mask_host):
            idx, buf = in_gpu.popleft()
            mask = mask_host.numpy().copy()           # the streamer reuses its pinned mask buffers
            writes.append(self.write_pool.submit(self.write, outs[idx], mask[..., 0] if mask.shape[-1] == 1 else mask))
            self._free.append(buf)

        feed()
        while pending:
            idx, buf, fut = pending.popleft()
            fut.result()
            nxt_buf = None
            if pending:
                pending[0][2].result()
                nxt_buf = pending[0][1]
            prev = self.streamer.submit(buf, nxt_buf)
            in_gpu.append((idx, buf))
            if prev is not None:
                retire(prev)
            feed()
        last = self.streamer.flush()
        while in_gpu:
            retire(last)
        for w in writes:
            w.result()
        return outs

    def close(self):
        self.decode_pool.shutdown(wait=True)
        self.write_pool.shutdown(wait=True)


def predict_tiled(image, model, test_transform, patch_size, batch_size, tile_step=None, tta=True, weight='pyramid'):
    """Reference signature (inria_submit.py:237) plus the knobs it hard-codes.  `image` is the raw uint8 array
    read_rgb returns; `test_transform` must be the reference-style normalisation (it is folded into a LUT)."""
    norm = find_normalize(test_transform)
    if norm is None:
        raise NotImplementedError("test_transform must be Sequential([ImageOnly(NormalizeImage(...))])")
    if image.dtype != np.uint8:
        raise NotImplementedError("the fused pipeline takes the uint8 image (normalisation happens on the device)")
    cache = model.__dict__.setdefault('_tiled_predictors', {})
    key = (image.shape, patch_size, tile_step, batch_size, tta, weight, tuple(norm.mean), tuple(norm.std), norm.scale,
           getattr(model, 'precision', 'bf16'), model._stamp())
    if key not in cache:
        cache.clear()
        cache[key] = TiledPredictor(model, image.shape, patch_size, tile_step, batch_size, weight, tta, norm)
    return cache[key](image)


def mask_from_probability(mask):
    """inria_submit.py:305."""
    return ((mask > 0.5) * 255).astype(np.uint8)
