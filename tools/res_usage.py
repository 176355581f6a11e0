"""Per-kernel resource usage of the built library (registers, static shared memory, local
memory = spills, stack), from ``cuobjdump -res-usage``:
    python tools/res_usage.py [out.md]        (needs cuobjdump + c++filt, no GPU)"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sup3r_b200", "lib", "libsup3r_b200.so")
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_res_usage.md")

txt = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
rows = []
for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", txt):
    use = dict(kv.split(":") for kv in m.group(2).split() if ":" in kv)
    rows.append((m.group(1), use))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True,
                       text=True).stdout.splitlines()
lines = ["# Resource usage per kernel of libsup3r_b200.so (sm_100a), round 2", "",
         "`python tools/res_usage.py` (`cuobjdump -res-usage` of the built library).  REG = registers "
         "per thread, SHARED = static shared memory (the tcgen05 kernels take their tiles as "
         "dynamic shared memory on top), LOCAL = local memory per thread (register spills / "
         "local arrays), STACK = call stack.", "",
         "| kernel | REG | SHARED | LOCAL | STACK |", "|---|---|---|---|---|"]
spill = 0
for (_, use), name in sorted(zip(rows, names), key=lambda t: -int(t[0][1].get("REG", 0))):
    name = re.sub(r"\(.*", "", name.replace("void ", ""))
    lines.append(f"| `{name}` | {use.get('REG')} | {use.get('SHARED')} | {use.get('LOCAL')} | "
                 f"{use.get('STACK')} |")
    spill += int(use.get("LOCAL", 0)) > 0 or int(use.get("STACK", 0)) > 0
lines += ["", f"{len(rows)} kernels; {spill} with a local-memory stack frame (register-pressure "
          "spills or locally indexed arrays)."]
# where the hot kernel's frame is touched: local loads / stores between consecutive tcgen05.mma
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
for tag, label in (("zring_kernelILi4ELi6E", "conv_umma_zring_kernel<4, 6>"),):
    inside, gaps, with_local, local, n_local = False, 0, 0, 0, 0
    seen = False
    for line in sass.splitlines():
        if "Function :" in line:
            inside = tag in line
            seen = False
            continue
        if not inside:
            continue
        if re.search(r"UTC[HQ]MMA", line):
            if seen:
                gaps += 1
                with_local += local > 0
            seen, local = True, 0
        elif re.search(r"\b(LDL|STL)\b", line):
            local += 1
            n_local += 1
    lines += ["", f"`{label}` (the body convolution): {n_local} static LDL / STL instructions; "
              f"{with_local} of the {gaps} gaps between consecutive `tcgen05.mma` instructions of "
              "the issuing thread contain one (the rest of the frame traffic is in the producer / "
              "epilogue code)."]
open(OUT, "w").write("\n".join(lines) + "\n")
print("wrote", OUT, len(rows), "kernels,", spill, "with a stack frame")
