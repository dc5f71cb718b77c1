"""Group an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total time, share.
Usage: python tools/launch_share.py launches.csv > share.md"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = re.sub(r"^void ", "", r[k])
    name = re.sub(r"\(.*$", "", name)[:78]
    tot[name] += float(r[v].replace(",", "")) / 1e3
    cnt[name] += 1
whole = sum(tot.values())
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for name, us in tot.most_common():
    if us / whole < 5e-4:
        continue
    print("| `%s` | %d | %.1f | %.1f %% |" % (name, cnt[name], us, 100 * us / whole))
conv = sum(us for n, us in tot.items() if "conv_" in n)
print("\nAll convolution kernels together: %.1f %% of the profiled time (%d launches, %.1f ms)." % (100 * conv / whole, sum(cnt.values()), whole / 1e3))
