"""Per-kernel device time of one generator step (warm, eager launches) via torch.profiler.
  python tools/step_breakdown.py [--chunks 8] [--precision bf16] [--detail]
"""
import argparse, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from sup3r_b200.models import Sup3rGan
from sup3r_b200 import configs as C

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=8)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--detail", action="store_true")
ap.add_argument("--out", default=None)
a = ap.parse_args()
dev = torch.device("cuda", 0)
hl = bench.gen_config()
Sup3rGan.seed(0)
model = Sup3rGan(hl, C.discriminator(3, "same", (2048, 1024)), precision=a.precision)
B = a.chunks
model.generator.build((B, *bench.LR_CHUNK))
x = torch.randn((B, *bench.LR_CHUNK), device=dev)
plan = model.plan_for(model.generator, a.precision)
for _ in range(3):
    plan.run(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        plan.run(x)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
n = len(evs) // 3
last = evs[-n:]
agg = {}
lines = []
for e in last:
    d = agg.setdefault(e.name[:70], [0, 0.0])
    d[0] += 1
    d[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    if a.detail:
        lines.append(f"{e.name[:60]:60s} {(e.device_time if hasattr(e,'device_time') else e.cuda_time):9.1f}")
tot = sum(v[1] for v in agg.values())
span = (last[-1].time_range.end - last[0].time_range.start)
out = [f"one step, {B} chunks, {a.precision}: {n} kernels, sum {tot:.1f} us, span {span:.1f} us"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k:70s} n={v[0]:3d} total={v[1]:9.1f} us  avg={v[1]/v[0]:8.1f}  {100*v[1]/tot:5.1f}%")
out += lines
txt = "\n".join(out)
print(txt)
if a.out:
    open(a.out, "w").write(txt + "\n")
