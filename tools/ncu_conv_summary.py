"""Summarise the `ncu --set full` capture of one UNet16 plan run (tools/layer_times.py 13) into profiles/rNN_ncu_conv_summary.{md,json}.
Usage: ncu -i X.ncu-rep --page raw --csv > X.csv; python tools/ncu_conv_summary.py X.csv profiles/r02_ncu_conv_summary [tiles]"""
import csv
import json
import sys

LAYERS = ["conv1_1 3->64 @512 (first layer, packed 3-channel tile)", "conv1_2 64->64 @512 (+pool)", "conv2_1 64->128 @256",
          "conv2_2 128->128 @256 (+pool)", "conv3_1 128->256 @128", "conv3_2 256->256 @128", "conv3_3 256->256 @128 (+pool)",
          "conv4_1 256->512 @64", "conv4_2 512->512 @64", "conv4_3 512->512 @64 (+pool)", "conv5_1 512->512 @32",
          "conv5_2 512->512 @32", "conv5_3 512->512 @32 (+pool)", "center conv 512->512 @16", "center ConvT 512->256 @16",
          "dec5 conv 768->512 @32", "dec5 ConvT 512->256 @32", "dec4 conv 768->512 @64", "dec4 ConvT 512->256 @64",
          "dec3 conv 512->256 @128", "dec3 ConvT 256->64 @128 (4 phases fused)", "dec2 conv 192->128 @256",
          "dec2 ConvT 128->32 @256 (4 phases fused, resident weights)",
          "dec1 conv 96->32 @512 + 1x1 head + sigmoid (resident weights)"]
# algorithmic GFLOP of the layer for 13 tiles of 512 x 512 (2 * pixels * Cin * Cout * taps; ConvT: 4 taps per OUTPUT pixel)
SHAPES = [(512, 3, 64, 9), (512, 64, 64, 9), (256, 64, 128, 9), (256, 128, 128, 9), (128, 128, 256, 9), (128, 256, 256, 9),
          (128, 256, 256, 9), (64, 256, 512, 9), (64, 512, 512, 9), (64, 512, 512, 9), (32, 512, 512, 9), (32, 512, 512, 9),
          (32, 512, 512, 9), (16, 512, 512, 9), (32, 512, 256, 4), (32, 768, 512, 9), (64, 512, 256, 4), (64, 768, 512, 9),
          (128, 512, 256, 4), (128, 512, 256, 9), (256, 256, 64, 4), (256, 192, 128, 9), (512, 128, 32, 4), (512, 96, 32, 9)]
TILES = int(sys.argv[3]) if len(sys.argv) > 3 else 13
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, name):
    return float(r[ix[name]].replace(",", "") or 0)


out = []
for layer, (hw, ci, co, taps), r in zip(LAYERS, SHAPES, rows[2:]):
    us = f(r, "gpu__time_duration.sum") * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
                                           "s": 1e6}[rows[1][ix["gpu__time_duration.sum"]]]
    rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    scale = {"Mbyte": 1.0, "Gbyte": 1e3, "Kbyte": 1e-3, "byte": 1e-6}
    rd *= scale[rows[1][ix["dram__bytes_read.sum"]]]
    wr *= scale[rows[1][ix["dram__bytes_write.sum"]]]
    gflop = 2.0 * TILES * hw * hw * ci * co * taps / 1e9
    out.append({"layer": layer, "kernel": r[ix["Kernel Name"]].replace("void ", "").split("(")[0], "us": us,
                "tflops": gflop / us * 1e3, "dram_read_MB": rd, "dram_write_MB": wr,
                "tensor_pipe_active_pct": f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
                if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in ix
                else f(r, "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"),
                "sm_ghz": f(r, "sm__cycles_elapsed.avg.per_second") if "sm__cycles_elapsed.avg.per_second" in ix else 0.0,
                "regs": int(f(r, "launch__registers_per_thread"))})
tot = sum(o["dram_read_MB"] + o["dram_write_MB"] for o in out)
tus = sum(o["us"] for o in out)
tfl = sum(o["tflops"] * o["us"] for o in out) / tus
js = {"source": "ncu --set full --clock-control none, tools/layer_times.py %d (UNet16, %d tiles of 512x512, bf16), the %d conv launches "
      "of one plan run" % (TILES, TILES, len(out)), "tile_batch": TILES, "total_dram_bytes_per_plan_run": tot * 1e6, "conv_launches": len(out),
      "mean_dram_bytes_per_launch": tot * 1e6 / len(out), "total_us": tus, "tflops_over_plan_run": tfl, "launches": out}
json.dump(js, open(sys.argv[2] + ".json", "w"), indent=1)
with open(sys.argv[2] + ".md", "w") as fo:
    fo.write("# ncu --set full, conv kernels of one UNet16 plan run (%d tiles of 512x512, bf16)\n\n" % TILES)
    fo.write("Command: `ncu --set full --clock-control none -k regex:\"conv_halo|conv_first|conv_igemm\" -s 48 -c 24 python tools/layer_times.py %d`" % TILES +
             " (B200).\n`tensor pipe` = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active; clocks are what the power cap "
             "allowed during the replay; TFLOP/s = algorithmic FLOPs / ncu duration (cold caches, serialised).\n\n")
    fo.write("| layer | kernel | time us | TFLOP/s | DRAM read MB | DRAM write MB | tensor pipe % | SM GHz | regs |\n|---|---|---:|---:|---:|---:|---:|---:|---:|\n")
    for o in out:
        fo.write("| %s | `%s` | %.1f | %.0f | %.1f | %.1f | %.1f | %.2f | %d |\n" % (
            o["layer"], o["kernel"], o["us"], o["tflops"], o["dram_read_MB"], o["dram_write_MB"], o["tensor_pipe_active_pct"],
            o["sm_ghz"] / 1e9 if o["sm_ghz"] > 1e6 else o["sm_ghz"], o["regs"]))
    fo.write("\nTotal DRAM traffic of the %d launches: %.1f MB (%.1f MB per launch on average); total time %.1f us; %.0f TFLOP/s over the "
             "plan run.\n" % (len(out), tot, tot / len(out), tus, tfl))
