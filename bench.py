#!/usr/bin/env python
"""Headline benchmark: megapixels/s of tiled UNet16 ("AlbuNet") inference on synthetic 5000x5000 Inria-shaped
images, 512 tile / 384 step, pyramid-weighted merge (BASELINE.json configs[2]; configs[3] when --gpus > 1).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores (oracle port)

A step = one image per rank through split -> UNet16 -> merge -> threshold -> confusion counts (+ NCCL all-reduce of
the counts and gather of the masks when N > 1).  Rank 0 prints ONE JSON line (see the driver contract in DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMAGE_HW = 5000
TILE, STEP = 512, 384
MPX_PER_IMAGE = IMAGE_HW * IMAGE_HW / 1e6
METRIC = "megapixels/sec tiled U-Net inference (UNet16/AlbuNet, 5000x5000, 512/384)"
# secondary workloads (--model): name -> (constructor, synthetic state_dict, tile, step, default tile batch, label)
MODELS = {
    # 85 tiles per launch: 169 = 85 + 84 (one padded slot), and the 64 x 64 / 128 x 128 layers then fill 36.8 / 73.5 waves of
    # 148 CTAs instead of 5.6 / 11.2 (measured: 13 -> 555, 57 -> 572, 85 -> 575, 169 -> 563 Mpx/s)
    "unet16": ("UNet16", lambda s: s.vgg_unet_state_dict("unet16", seed=0), 512, 384, 85, "configs[2]: UNet16"),
    "unet11": ("UNet11", lambda s: s.vgg_unet_state_dict("unet11", seed=0), 512, 384, 85, "UNet11 (TernausNet-VGG11)"),
    # 1936 tiles of 224 x 224: 11 launches of 176 (zf_unet: 44 -> 372, 176 -> 395 Mpx/s), 16 launches of 121 (fcdensenet67)
    "zf_unet": ("ZF_UNET", lambda s: s.zf_unet_state_dict(seed=0), 224, 112, 176, "ZF_UNET"),
    "linknet34": ("LinkNet34", lambda s: s.linknet34_state_dict(seed=0), 512, 384, 169, "LinkNet34 (configs[1] model, eval forward)"),
    "fcdensenet67": ("FCDenseNet67", lambda s: s.fcdensenet_state_dict(seed=0), 224, 112, 121, "configs[4]: FCDenseNet67"),
}


def bench_config(n_gpus, tta=False):
    """`config` of the JSON line: identical keys and values in the CUDA arm and the reference arm (the driver compares
    them); everything specific to one arm goes into `run`."""
    workload = "configs[2]: UNet16 tiled inference, 5000x5000x3 u8, tile 512 / step 384, pyramid merge"
    if n_gpus > 1:
        workload += "; configs[3]: one image per rank per step, NCCL all-reduce of IoU counts + gather of masks"
    return {"workload": workload, "tiles_per_image": 169, "tta": bool(tta)}


FLOP_PER_TILE_UNET16 = 319689850880            # 25 convolutions + the 1x1 head of one 512 x 512 tile, no padding
FLOP_PER_TILE_FCDENSENET67 = 41514292736       # one 224 x 224 tile
FLOP_PER_SAMPLE_LINKNET34 = 11.64e9            # forward of one 256 x 256 sample (x3 with backward)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p.get("bf16_tflops_sustained", 1400.0)), float(p.get("hbm_gbs", 6650.0)), "measured"
    return 1400.0, 6650.0, "fallback"   # B200_PROFILING.md: ~1.4 PFLOP/s sustained, 6.65 TB/s


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (NVML, 20 ms period; nvidia-smi fallback)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()
        self.max_mhz, self.nvml, self.handle = None, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if index < len(ids) and ids[index].strip().isdigit():
                return int(ids[index])
        return index

    def _sample_nvml(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(
            n, "nvmlDeviceGetCurrentClocksEventReasons") else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        flags = [bool(r & 0x8), bool(r & 0x40), bool(r & 0x20), bool(r & 0x4)]   # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
        return mhz, flags

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        f = [v.strip() for v in out.strip().split(",")]
        self.max_mhz = float(f[1])
        return float(f[0]), [v.lower().startswith("active") for v in f[2:6]]

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self.samples.append(self._sample_nvml() if self.nvml else self._sample_smi())
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nvml else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        mhz = sorted(s[0] for s in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[1][i] for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_min_mhz": mhz[0], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(mhz), "source": "nvml" if self.nvml else "nvidia-smi"}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference_step(sd, image, n_net_tiles):
    """The reference algorithm (oracle port) on the host: float64 normalise + split of the whole image, UNet16 fp32
    on `n_net_tiles` tiles (extrapolated to all 169), pyramid merge of 169 tiles, threshold.  Returns seconds/image."""
    from oracle import nets_oracle as no
    from oracle import tiles_oracle as to

    t0 = time.perf_counter()
    x = to.normalize_image(image)
    s = to.SlicerOracle(x.shape, TILE, STEP, weight="pyramid")
    tiles = s.split(x)
    t1 = time.perf_counter()
    with torch.no_grad():
        batch = torch.from_numpy(to.to_nchw_float(tiles[:n_net_tiles]))
        probs = torch.sigmoid(no.unet_vgg_forward(sd, batch, "unet16")).numpy()
    t2 = time.perf_counter()
    preds = [np.moveaxis(probs[i % n_net_tiles], 0, -1) for i in range(len(tiles))]
    mask = ((s.merge(preds, dtype=np.float32) > 0.5) * 255).astype(np.uint8)
    t3 = time.perf_counter()
    assert mask.shape == (IMAGE_HW, IMAGE_HW, 1)
    return (t1 - t0) + (t2 - t1) * len(tiles) / n_net_tiles + (t3 - t2)


def cpu_baseline(n_net_tiles=4, repeats=1):
    from snb_b200 import synth

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.vgg_unet_state_dict("unet16", seed=0)
    image = synth.image_u8(0, IMAGE_HW, IMAGE_HW)
    secs = [cpu_reference_step(sd, image, n_net_tiles) for _ in range(repeats)]
    return {"value": MPX_PER_IMAGE / min(secs), "unit": "Mpx/s", "cores": cores, "kind": "port",
            "sample": "full float64 normalise+split and pyramid merge of one 5000x5000 image; UNet16 fp32 (CPU PyTorch) "
                      "on %d of 169 tiles, net time extrapolated x169/%d" % (n_net_tiles, n_net_tiles)}


def run_reference(args):
    """Reference arm: the reference algorithm (oracle port) on the host cores.  One full float64 normalise + split and one
    pyramid merge of a 5000x5000 image are timed once (they do not depend on the step); every timed step then runs UNet16
    fp32 on `n_net` of the 169 tiles and is extrapolated x169/n_net.  n_net is chosen so that the whole --steps K --warmup W
    run stays within SNB_REF_BUDGET_S seconds (default 150) whatever K and W the driver passes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import nets_oracle as no
    from snb_b200 import synth
    from oracle import tiles_oracle as to

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.vgg_unet_state_dict("unet16", seed=0)
    image = synth.image_u8(0, IMAGE_HW, IMAGE_HW)
    budget = float(os.environ.get("SNB_REF_BUDGET_S", "150"))

    t0 = time.perf_counter()
    x = to.normalize_image(image)
    slicer = to.SlicerOracle(x.shape, TILE, STEP, weight="pyramid")
    tiles = slicer.split(x)
    t_split = time.perf_counter() - t0

    def net(n_tiles):
        t = time.perf_counter()
        with torch.no_grad():
            batch = torch.from_numpy(to.to_nchw_float(tiles[:n_tiles]))
            probs = torch.sigmoid(no.unet_vgg_forward(sd, batch, "unet16")).numpy()
        return time.perf_counter() - t, probs

    t_one, probs = net(1)                                              # also the warm-up of the ATen kernels
    preds = [np.moveaxis(probs[0], 0, -1) for _ in range(len(tiles))]
    t0 = time.perf_counter()
    mask = ((slicer.merge(preds, dtype=np.float32) > 0.5) * 255).astype(np.uint8)
    t_merge = time.perf_counter() - t0
    assert mask.shape == (IMAGE_HW, IMAGE_HW, 1)
    n_steps = max(1, args.steps + args.warmup)
    n_net = int(max(1, min(4, (budget - t_split - t_merge) / (n_steps * max(t_one, 1e-3)))))
    for _ in range(args.warmup):
        net(n_net)
    secs = []
    for _ in range(args.steps):
        t_net, _ = net(n_net)
        secs.append(t_split + t_net * len(tiles) / n_net + t_merge)
    sec = sum(secs) / len(secs)
    value = MPX_PER_IMAGE / sec
    sample = ("full float64 normalise+split (%.1f s) and pyramid merge (%.1f s) of one 5000x5000 image timed once; per step "
              "UNet16 fp32 on %d of 169 tiles extrapolated x169/%d (oracle port of lib/tiles.py + lib/models/unet16.py; the "
              "reference is Python and cannot travel to the GPU box)" % (t_split, t_merge, n_net, n_net))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpx/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.gpus, args.tta),
            "run": {"extrapolated": True, "net_tiles_timed_per_step": n_net, "tiles_per_image": len(tiles),
                    "note": "ms_per_step is an EXTRAPOLATION: split and merge of one whole image timed once, the network timed "
                            "on %d of 169 tiles per step and scaled x169/%d; a full image takes minutes on the host" % (n_net, n_net)},
            "cpu_baseline": {"value": value, "unit": "Mpx/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ CUDA arm
def _fence(world):
    import torch.distributed as dist

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _timed(fn, steps, world, dev, after=None):
    """K calls of fn bracketed by barrier + synchronize on both sides, CUDA events on the launching stream, MAX over ranks."""
    import torch.distributed as dist

    _fence(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    if after is not None:
        after()
    e1.record()
    _fence(world)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def _mask_checksum(mask):
    """position-weighted 64-bit checksum of a uint8 mask (a permutation or offsetting errors change it)"""
    v = mask.reshape(-1).to(torch.int64)
    w = (torch.arange(v.numel(), device=v.device, dtype=torch.int64) % 65521) + 1
    return (v * w).sum()


def run_cuda(args):
    import torch.distributed as dist

    import snb_b200  # noqa: F401
    from snb_b200 import synth
    from snb_b200 import dist as sdist
    from snb_b200 import inria_submit as sub
    from snb_b200.engine import ConvOp
    from snb_b200.lib import metrics
    rank, world, local_rank = sdist.init_from_env()
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    from snb_b200.lib import models as M

    cls_name, make_sd, tile, step, default_batch, label = MODELS[args.model]
    tile, step = args.tile or tile, args.step or step
    cls = getattr(M, cls_name)
    model = cls(n_classes=1) if args.model == "fcdensenet67" else cls()
    model.load_state_dict(make_sd(synth))
    model = model.to(dev).eval()
    if args.precision != "bf16":
        model.set_precision(args.precision)
    headline = args.model == "unet16" and (tile, step) == (TILE, STEP) and args.precision == "bf16" and not args.tta
    if args.shard == "tile":
        line = run_tile_sharded(args, model, tile, step, dev, rank, world, label, cls_name)
        if rank == 0:
            print(json.dumps(line), flush=True)
        return
    pred = sub.TiledPredictor(model, (IMAGE_HW, IMAGE_HW, 3), tile, step, batch_size=args.batch or default_batch,
                              tta=args.tta, device=dev, use_graph=args.graph)
    metric = METRIC if headline else ("megapixels/sec tiled inference (%s, 5000x5000, %d/%d%s%s)" % (
        cls_name, tile, step, ", D4 TTA" if args.tta else "", ", " + args.precision if args.precision != "bf16" else ""))

    # distinct synthetic images per rank (weak scaling: one image per rank per step)
    n_img = 2
    host_imgs = [torch.from_numpy(synth.image_u8(100 * rank + i, IMAGE_HW, IMAGE_HW)).pin_memory() for i in range(n_img)]
    dev_imgs = [h.to(dev) for h in host_imgs]
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    gts = [(torch.rand((IMAGE_HW, IMAGE_HW, 1), device=dev, generator=g) > 0.5).to(torch.uint8) for _ in range(n_img)]
    host_counts = torch.empty(4, dtype=torch.int64).pin_memory()
    counts_buf = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(2)]
    # the only exchange of the image-sharded job: all-reduce of int64[4] + gather of u8 masks, on a side stream with
    # preallocated double buffers so that it overlaps the next image's convolutions
    exchange = sdist.MaskExchange((IMAGE_HW, IMAGE_HW, 1), dev) if world > 1 else None
    last_slot = [0]

    def finish_image(i, merged, mask):
        counts = metrics.confusion_counts_from_probs(merged, gts[i % n_img], out=counts_buf[i & 1])
        if exchange is not None:
            last_slot[0] = exchange.submit(mask, counts)
        return counts

    def step_resident(i):
        merged, mask = pred.predict_device(dev_imgs[i % n_img])
        return finish_image(i, merged, mask)

    streamer = sub.StreamingPredictor(pred)

    def step_e2e(i):
        # H2D of this step's image from pinned memory and D2H of its mask both happen inside the timed region, on a
        # copy stream that overlaps the neighbouring images' compute (the public host-to-host API)
        streamer.submit(host_imgs[i % n_img], host_imgs[(i + 1) % n_img])
        counts = finish_image(i, pred.merged, pred.mask)
        if exchange is not None:
            exchange.wait()
            counts = exchange.result(last_slot[0])[0]
        host_counts.copy_(counts, non_blocking=True)

    drain = exchange.wait if exchange is not None else None
    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total = _timed(step_resident, args.steps, world, dev, after=drain)     # the timed region: one CUDA-graph replay per image
    clocks = sampler.summary()

    # multi-rank integrity: what rank 0 gathered for rank r is byte-for-byte rank r's mask (position-weighted checksums)
    gathered_ok = None
    if world > 1:
        exchange.wait()
        mine = _mask_checksum(pred.mask).reshape(1)
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        if rank == 0:
            g_masks = exchange.result(last_slot[0])[1]
            gathered_ok = all(int(_mask_checksum(g_masks[r])) == int(sums[r]) for r in range(world))

    # Per-launch conv timing with CUDA events.  Events cannot be recorded between the nodes of a replayed graph, so
    # the same K steps are run once more eagerly, right here, with one event after every launch of every plan run.
    marks = []
    plan = pred.plan
    orig_run = plan.run

    def run_marked():
        st = snb_b200._native.stream_ptr()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(plan.ops) + 1)]
        ev[0].record()
        for k, op in enumerate(plan.ops):
            op(st)
            ev[k + 1].record()
        marks.append(ev)
        return plan.out

    plan.run, pred.use_graph = run_marked, False
    ms_eager = _timed(step_resident, args.steps, world, dev, after=drain)
    plan.run, pred.use_graph = orig_run, args.graph

    conv_ms, conv_flops, conv_launches, other_ms = 0.0, 0.0, 0, 0.0
    for ev in marks:
        for k, op in enumerate(plan.ops):
            dt = ev[k].elapsed_time(ev[k + 1])
            if isinstance(op, ConvOp):
                conv_ms += dt
                conv_flops += op.flops
                conv_launches += 1
            else:
                other_ms += dt

    for i in range(min(2, args.warmup)):
        step_e2e(i)
    streamer.flush()

    def e2e_steps(i):
        step_e2e(i)
        if i == args.steps - 1:
            streamer.flush()                                                 # the last mask lands on the host inside the region

    ms_e2e = _timed(e2e_steps, args.steps, world, dev, after=drain)

    ms_step = ms_total / args.steps
    value = world * MPX_PER_IMAGE / (ms_step / 1e3)
    e2e_value = world * MPX_PER_IMAGE / (ms_e2e / args.steps / 1e3)
    peak_tf, peak_bw, peak_src = measured_peaks()
    achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    # DRAM bytes per conv launch from the committed ncu --set full capture of the same plan (profiles/), headline config only
    traffic, traffic_src = None, None
    for name in ("r02_ncu_conv_summary.json", "r01_ncu_conv_summary.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if headline and os.path.exists(tpath):
            with open(tpath) as fh:
                summ = json.load(fh)
            if int(summ.get("tile_batch", 13)) == pred.batch:   # per-launch bytes only compare at the same tiles per launch
                traffic, traffic_src = float(summ["mean_dram_bytes_per_launch"]), "profiles/" + name
                break
    line = {
        "metric": metric, "value": value, "unit": "Mpx/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
        "data": "synthetic",
        "config": bench_config(world, args.tta) if headline else {
            "workload": "%s tiled inference, 5000x5000x3 u8, tile %d / step %d, pyramid merge" % (label, tile, step),
            "tiles_per_image": pred.n_tiles, "tta": bool(args.tta)},
        "run": {"tile_batch": pred.batch, "cuda_graph": bool(args.graph),
                "l2_policy": "no flush needed: per-step working set (activations of %d tiles/batch, GBs) >> 126 MB L2; "
                             "%d images rotate" % (pred.batch, n_img),
                "flop_per_image": pred.flops_per_image,
                "flop_per_image_survey_8d": 169 * FLOP_PER_TILE_UNET16 if headline else None,
                "exchange": None if world == 1 else "all-reduce int64[4] + gather u8 masks on a side stream (double-buffered)",
                "gathered_masks_match_rank_masks": gathered_ok},
        "e2e": {"value": e2e_value, "unit": "Mpx/s", "h2d_bytes_per_step": IMAGE_HW * IMAGE_HW * 3,
                "d2h_bytes_per_step": IMAGE_HW * IMAGE_HW + 32},
        "gpu_launches": (pred.launches_per_image + 1) * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf, "traffic": traffic,
                     "traffic_note": "mean dram__bytes_read+write per conv launch of one plan run at the same tiles per launch (ncu, %s); "
                                     "algorithmic FLOPs per launch vary per layer" % traffic_src,
                     "kernel": "conv_halo_kernel / conv_first_kernel (tcgen05 implicit GEMM, all %d conv launches of the timed region)" % conv_launches,
                     "peak_source": peak_src + " bf16_tflops_sustained",
                     "timing": "CUDA events after every launch in an eager re-run of the same K steps (the timed "
                               "region itself replays a CUDA graph); shares are of that eager pass",
                     "conv_share_of_step": conv_ms / ms_eager, "conv_ms_per_step": conv_ms / args.steps,
                     "eager_ms_per_step": ms_eager / args.steps,
                     "other_plan_ms_per_step": other_ms / args.steps,
                     "whole_step_tflops": pred.flops_per_image / (ms_step / 1e3) / 1e12},
    }
    secondary = {}
    if headline and not args.no_secondary:
        # free the headline's buffers before the secondary workloads allocate theirs
        del streamer, pred, plan, marks, dev_imgs, gts
        model._plans = {}
        torch.cuda.empty_cache()
        if world > 1:
            sargs = argparse.Namespace(**vars(args))
            secondary["tile_sharded"] = run_tile_sharded(sargs, model, tile, step, dev, rank, world, label, cls_name, brief=True)
        else:
            secondary = run_secondary(args, dev, peak_tf)
            line["roofline_hbm"] = run_hbm(peak_bw, dev)
    if secondary:
        line["secondary"] = secondary
    if rank != 0:
        return
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line), flush=True)


def run_hbm(peak_bw, dev):
    """HBM-bound kernels of the path at BASELINE sizes: algorithmic bytes / CUDA-event time / measured copy bandwidth."""
    from snb_b200 import hbm_bench

    try:
        r = hbm_bench.measure(peak_bw, reps=5, device=dev)
        r["_how"] = ("CUDA events around 5 back-to-back calls; every call moves > 126 MB and the inputs rotate over 3-4 sets "
                     "(> 2 x L2), so no call finds its inputs in L2; bytes are the algorithmic bytes of SURVEY 8d; frac is "
                     "against the measured copy bandwidth (%.0f GB/s)" % peak_bw)
        return r
    except Exception as exc:                          # a secondary measurement must not take the headline line down
        return {"error": repr(exc)}
    finally:
        torch.cuda.empty_cache()


def run_secondary(args, dev, peak_tf):
    """configs[4] (FCDenseNet67 tiled inference) and configs[1] (LinkNet34 forward / backward step) on this GPU."""
    from snb_b200 import synth
    from snb_b200 import inria_submit as sub
    from snb_b200.lib import losses
    from snb_b200.lib import models as M

    out = {}
    try:
        m = M.FCDenseNet67(n_classes=1)
        m.load_state_dict(synth.fcdensenet_state_dict(seed=0))
        m = m.to(dev).eval()
        pred = sub.TiledPredictor(m, (IMAGE_HW, IMAGE_HW, 3), 224, 112, batch_size=MODELS["fcdensenet67"][4], tta=False, device=dev)
        img = torch.from_numpy(synth.image_u8(0, IMAGE_HW, IMAGE_HW)).to(dev)
        steps = 2
        for _ in range(2):
            pred.predict_device(img)
        ms = _timed(lambda i: pred.predict_device(img), steps, 1, dev) / steps
        mpx = MPX_PER_IMAGE / (ms / 1e3)
        roof = MPX_PER_IMAGE / (pred.n_tiles * FLOP_PER_TILE_FCDENSENET67 / (peak_tf * 1e12))
        out["fcdensenet67_configs4"] = {"metric": "megapixels/sec tiled inference (FCDenseNet67, 5000x5000, 224/112, bf16)",
                                        "value": mpx, "unit": "Mpx/s", "ms_per_image": ms, "tiles_per_image": pred.n_tiles,
                                        "tile_batch": pred.batch, "steps": steps,
                                        "roofline_mpx": roof, "frac": mpx / roof,
                                        "note": "roofline = 1936 tiles x 41.51 GFLOP (SURVEY 8d) at the measured sustained bf16 peak"}
        del pred, m, img
    except Exception as exc:
        out["fcdensenet67_configs4"] = {"error": repr(exc)}
    torch.cuda.empty_cache()
    try:
        batch, size = 8, 256
        m = M.LinkNet34(pretrained=False)
        m.load_state_dict(synth.linknet34_state_dict(seed=0))
        m = m.to(dev).train()
        rs = np.random.RandomState(0)
        x = torch.from_numpy(rs.standard_normal((batch, 3, size, size)).astype(np.float32)).to(dev)
        t = torch.from_numpy((rs.rand(batch, 1, size, size) > 0.5).astype(np.int64)).to(dev)
        crit = losses.BCEWithLogitsLossAndSmoothJaccard()
        for _ in range(3):
            m.train_step(x, t, crit)
        steps = 20
        ms = _timed(lambda i: m.train_step(x, t, crit), steps, 1, dev) / steps
        with torch.no_grad():
            ms_fwd = _timed(lambda i: m(x), steps, 1, dev) / steps
        plan = m.plan_train(batch, size, size)
        flop = 3 * batch * FLOP_PER_SAMPLE_LINKNET34
        out["linknet34_train_step_configs1"] = {
            "metric": "LinkNet34 forward + bce_jaccard + backward, batch 8 x 3 x 256 x 256, bf16 (Dropout2d p = 0.5 active)",
            "ms_per_step": ms, "images_per_s": batch / ms * 1e3, "forward_only_ms": ms_fwd, "steps": steps,
            "gpu_launches_per_step": plan.launches + plan.bwd_launches + 2,
            "algorithmic_tflops": flop / (ms / 1e3) / 1e12,
            "note": "model.train_step: forward graph + fused loss + backward graph (tcgen05 dgrad / wgrad), no optimiser; "
                    "algorithmic FLOPs = 3 x 8 x 11.64 GFLOP (SURVEY 8a); the step is launch- and BatchNorm-bound, not tensor-bound"}
        del m, plan
    except Exception as exc:
        out["linknet34_train_step_configs1"] = {"error": repr(exc)}
    torch.cuda.empty_cache()
    return out


def run_tile_sharded(args, model, tile, step, dev, rank, world, label, cls_name, brief=False):
    """ONE image per step split by crop range over all ranks (BASELINE configs[3] "sharded by tile", strong scaling of
    single-image latency): seam tiles by NCCL send / recv, band merge per rank, all-gather of the uint8 mask bands and
    all-reduce of the counts on a side stream that overlaps the next image.  Every rank's mask is compared BYTE FOR BYTE
    with the single-GPU mask of the same image."""
    import torch.distributed as dist

    from snb_b200 import synth
    from snb_b200 import inria_submit as sub
    from snb_b200.lib import metrics

    pred = sub.TileShardedPredictor(model, (IMAGE_HW, IMAGE_HW, 3), tile, step, batch_size=args.batch or None, tta=args.tta,
                                    device=dev, use_graph=args.graph, overlap=True)
    imgs = [torch.from_numpy(synth.image_u8(i, IMAGE_HW, IMAGE_HW)).to(dev) for i in range(2)]
    g = torch.Generator(device=dev).manual_seed(1000)
    gt = (torch.rand((IMAGE_HW, IMAGE_HW, 1), device=dev, generator=g) > 0.5).to(torch.uint8)
    steps, warmup = args.steps, max(3, args.warmup)

    def step_fn(i):
        pred.predict_device(imgs[i % 2], gt)

    for i in range(warmup):
        step_fn(i)
    pred.wait()
    ms = _timed(step_fn, steps, world, dev, after=pred.wait)
    # byte equality with the single-GPU result of the last image, on every rank
    last = (steps - 1) % 2
    mask_sharded, counts_sharded = pred.local.mask.clone(), pred.counts.clone()
    tiles_per_rank = [e - b for b, e in pred.ranges]
    seam_bytes, tile_batch = pred.exchange_bytes, pred.local.batch
    del pred
    torch.cuda.empty_cache()
    solo = sub.TiledPredictor(model, (IMAGE_HW, IMAGE_HW, 3), tile, step, batch_size=13, tta=args.tta, device=dev, use_graph=False)
    merged, mask = solo.predict_device(imgs[last])
    counts = metrics.confusion_counts_from_probs(merged, gt)
    same = torch.tensor([1 if (torch.equal(mask, mask_sharded) and counts.tolist() == counts_sharded.tolist()) else 0], device=dev)
    if world > 1:
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
    ms_step = ms / steps
    line = {
        "metric": "megapixels/sec tiled inference, ONE image sharded by tile (%s, 5000x5000, %d/%d)" % (cls_name, tile, step),
        "value": MPX_PER_IMAGE / (ms_step / 1e3), "unit": "Mpx/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "%s, one image per step split by crop range over %d ranks; seam tiles by NCCL send/recv, band "
                               "merge per rank, all-gather of u8 mask bands, all-reduce of counts" % (label, world),
                   "tiles_per_image": 169, "tta": bool(args.tta)},
        "run": {"tiles_per_rank": tiles_per_rank, "tile_batch": tile_batch, "seam_bytes_received_per_image_rank0": seam_bytes,
                "exchange": "side stream, double-buffered tile store: overlaps the next image's convolutions"},
        "masks_and_counts_byte_identical_to_single_gpu": bool(int(same) == 1), "counts": counts.tolist()}
    del solo
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="tiles per network launch (default: per model)")
    ap.add_argument("--model", default="unet16", choices=sorted(MODELS), help="secondary workloads; the headline is unet16")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"],
                    help="bf16 (headline) or tf32 = fp32 storage with TF32 tensor-core products (the 1e-4 'fp32 mode')")
    ap.add_argument("--shard", default="image", choices=["image", "tile"],
                    help="multi-GPU partitioning: one image per rank per step (weak, default) or one image split by tile (strong)")
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--step", type=int, default=0)
    ap.add_argument("--tta", action="store_true", help="D4 test-time augmentation (8 views per tile)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads / HBM-kernel block of the headline run")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
        run_cuda(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
