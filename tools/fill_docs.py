"""Fill the @PLACEHOLDER@ numbers of DESIGN.md / README.md from the measurement artefacts in profiles/ (run once, after
tools/gpu_final.sh and tools/gpu_multi.sh; the placeholders are consumed)."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda n: os.path.join(ROOT, "profiles", n)
b = json.load(open(P("r02_bench_final.json")))
hbm = b["roofline_hbm"]
order = ["merge", "loss_i64", "counts", "pr_curve", "split_hwc", "split_norm", "loss_u8"]
hbm_s = " / ".join("%.2f" % hbm[k]["frac"] for k in order)
sec = {}
for line in open(P("r02_secondary.txt")):
    m = re.search(r"\((\w+),.*\| ([0-9.]+) Mpx/s", line)
    if m:
        sec[m.group(1)] = float(m.group(2))
    elif line.startswith("UNet16 with D4 TTA"):
        sec["tta"] = float(re.search(r"\| ([0-9.]+) Mpx/s", line).group(1))
    elif line.startswith("UNet16 tf32"):
        sec["tf32"] = float(re.search(r"\| ([0-9.]+) Mpx/s", line).group(1))
n2 = json.load(open(P("r02_bench_n2.json")))
t2 = json.load(open(P("r02_bench_tile_n2.json")))
n2_s = "%.0f / %.0f Mpx/s (%.1f ms per image tile-sharded)" % (n2["value"], t2["value"], t2["ms_per_step"])
if os.path.exists(P("r02_bench_n8.json")) and os.path.getsize(P("r02_bench_n8.json")) > 10:
    n8 = json.load(open(P("r02_bench_n8.json")))
    t8 = json.load(open(P("r02_bench_tile_n8.json")))
    n2_s += "; 8 GPUs: %.0f / %.0f Mpx/s (%.1f ms per image)" % (n8["value"], t8["value"], t8["ms_per_step"])
rep = {
    "@HEAD@": "%.0f" % b["value"], "@MS@": "%.1f" % b["ms_per_step"], "@E2E@": "%.0f Mpx/s" % b["e2e"]["value"],
    "@FRAC@": "%.2f" % b["roofline"]["frac"], "@TF@": "%.0f" % b["roofline"]["achieved"], "@HBM@": hbm_s, "@N2@": n2_s,
    "@SEC@": "%.0f / %.0f / %.0f / %.0f Mpx/s" % (sec.get("LinkNet34", 0), sec.get("UNet11", 0), sec.get("ZF_UNET", 0), sec.get("FCDenseNet67", 0)),
}
tr = b["secondary"]["linknet34_train_step_configs1"]
readme = ("%s Mpx/s for the headline config (%s ms per 5000x5000 image under the 1 kW power cap, SM clock %d MHz; conv kernels %s TFLOP/s = %s of\n"
          "the measured sustained cuBLAS bf16 peak), end to end from pinned host memory %s; box-to-box spread of the pool ~4 %%.\n"
          "2 GPUs (weak / one image tile-sharded): %s.  CPU oracle port %.2f Mpx/s on the box's %d host cores.\n"
          "HBM kernels as a fraction of the measured copy bandwidth (merge / loss int64 / counts / PR curve / split_hwc / split+normalise /\n"
          "loss u8): %s.  LinkNet34 (eval) %.0f, UNet11 %.0f, ZF_UNET %.0f, FCDenseNet67 %.0f Mpx/s; UNet16 with D4 TTA %.0f, tf32 mode %.0f.\n"
          "LinkNet34 training step (train-mode forward with Dropout2d, bce_jaccard, backward; batch 8 x 256^2): %.2f ms (forward %.2f ms),\n"
          "all convolutions, input and weight gradients on tcgen05.") % (
    rep["@HEAD@"], rep["@MS@"], b["clocks"]["sm_mhz"], rep["@TF@"], rep["@FRAC@"], rep["@E2E@"], n2_s, b["cpu_baseline"]["value"],
    b["cpu_baseline"]["cores"], hbm_s, sec.get("LinkNet34", 0), sec.get("UNet11", 0), sec.get("ZF_UNET", 0), sec.get("FCDenseNet67", 0),
    sec.get("tta", 0), sec.get("tf32", 0), tr["ms_per_step"], tr["forward_only_ms"])
rep["@README_NUMBERS@"] = readme
for name in ("DESIGN.md", "README.md"):
    path = os.path.join(ROOT, name)
    s = open(path).read()
    for k, v in rep.items():
        s = s.replace(k, v)
    open(path, "w").write(s)
    left = re.findall(r"@[A-Z_0-9]+@", s)
    print(name, "placeholders left:", left)
