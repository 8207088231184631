"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into one row per kernel.

    python scripts/aggregate_launches.py gpurun_out/r02_launches_c4_default.csv "header comment" > profiles/...csv
"""
import csv, re, sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as fh:
    lines = [ln for ln in fh if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in rd:
    name = r[ik]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)                      # drop the parameter list
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    name = re.sub(r"\((bool|int)\)", "", name)
    lib = name.startswith("sc::") or name.startswith("cub::")
    if not lib:
        name = "torch: " + name[:60]
    a = agg.setdefault(name, [0, 0.0, lib])
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) / 1e6
lib_total = sum(a[1] for a in agg.values() if a[2])
print("# " + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# share_pct is of the library's own kernels (sc:: and the CUB sorts / scans it calls); torch kernels are the synthetic data generators, result plumbing and the legs' FP64 checks")
print("kernel,launches,total_ms,share_pct,avg_ms")
for name, (cnt, ms, lib) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    share = f"{100 * ms / lib_total:.3f}" if lib else "nan"
    print(f'"{name}",{cnt},{ms:.3f},{share},{ms / cnt:.4f}')
