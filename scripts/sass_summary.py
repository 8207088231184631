"""Per-kernel SASS opcode counts of libsc_b200.so (evidence that the build is Blackwell-native):
    python scripts/sass_summary.py > profiles/r02_sass_opcodes.txt
Runs `cuobjdump -sass` on the built library and counts, per kernel, the mnemonics that identify tcgen05 / TMEM
(UTCHMMA, UTCBAR, LDTM, UTCATOMSWS...), TMA (UTMALDG, UBLKCP), mbarrier (SYNCS), cp.async (LDGSTS), packed FP32
(FADD2 / FFMA2), register reallocation (USETMAXREG) and the FP64 accumulation (DFMA / DADD)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "spatialcore_b200", "libsc_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], check=True, capture_output=True, text=True).stdout
WANT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "ARRIVES",
        "FADD2", "FFMA2", "FMUL2", "USETMAXREG", "DFMA", "DADD", "LDS", "STS", "LDG", "STG", "ATOM", "RED", "SHFL", "BAR", "HMMA", "IMMA"]
kern = None
counts = collections.OrderedDict()
arch = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*", "", kern)[:90]
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1).split(".")[0]
        counts[kern]["_total"] += 1
        for w in WANT:
            if op == w or (w in ("LDS", "STS", "LDG", "STG", "ATOM", "RED", "BAR") and op.startswith(w)):
                counts[kern][w] += 1
print(f"# cuobjdump -sass spatialcore_b200/libsc_b200.so  (arch {arch}); opcode counts per kernel, zero columns omitted")
tot = collections.Counter()
for k, c in counts.items():
    row = " ".join(f"{w}={c[w]}" for w in WANT if c[w])
    print(f"{k:<92s} insts={c['_total']:<6d} {row}")
    tot.update(c)
print("# library totals: " + " ".join(f"{w}={tot[w]}" for w in WANT if tot[w]) + f" insts={tot['_total']}")
