"""SASS mnemonic summary per kernel of the built library (static instruction counts):
    python tools/sass_summary.py [out.md]        (needs cuobjdump + c++filt, no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sup3r_b200", "lib", "libsup3r_b200.so")
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.md")
MN = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "HMMA", "FFMA", "SYNCS"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line) if cur else None
    if m:
        counts[cur]["_total"] += 1
        for k in MN:
            if m.group(1).startswith(k):
                counts[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True,
                       text=True).stdout.splitlines()
rows = []
for (f, c), name in zip(counts.items(), names):
    name = name.replace("void ", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\((?:s3::|CUtensorMap|float|unsigned|int|long|void|const).*$", "", name)
    rows.append((name, c))
out = ["# SASS mnemonic summary of libsup3r_b200.so (sm_100a), round 2", "",
       "`python tools/sass_summary.py` (cuobjdump -sass of the built library; static instruction",
       "counts per kernel).  UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = tcgen05.mma kind::f8f6f4,",
       "UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld,",
       "HMMA = legacy mma.sync, FFMA = fp32 FMA, SYNCS = mbarrier operations.", "",
       "| kernel | instr | " + " | ".join(MN) + " |", "|---|---|" + "---|" * len(MN)]
key = lambda r: -(r[1]["UTCHMMA"] + r[1]["UTCQMMA"]) * 1000 - r[1]["HMMA"] * 10 - r[1]["FFMA"] / 100
for name, c in sorted(rows, key=key):
    if c["UTCHMMA"] + c["UTCQMMA"] + c["HMMA"] + c["UTMALDG"] == 0 and c["FFMA"] < 200:
        continue
    out.append(f"| `{name[:80]}` | {c['_total']} | " + " | ".join(str(c[k]) for k in MN) + " |")
tot = collections.Counter()
for _, c in rows:
    tot.update(c)
out += ["", "Library totals over %d kernels: " % len(rows)
        + ", ".join(f"{tot[k]} {k}" for k in MN) + "."]
open(OUT, "w").write("\n".join(out) + "\n")
print("\n".join(out[-12:]))
