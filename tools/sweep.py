"""BASELINE configs[4]: large-tile inference sweep of the north-star generator -- LR chunk
8x8x12 ... 64x64x96, batch 1..32 -- device-timed (CUDA events, graph replay, L2 flushed), with the
algorithmic TFLOP/s and the fraction of the measured sustained bf16 peak.
  python tools/sweep.py [--out profiles/r01_sweep.md]"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from sup3r_b200.models import Sup3rGan
from sup3r_b200 import configs as C

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--precision", default="fp16c")
ap.add_argument("--max-hr-gb", type=float, default=24.0, help="skip cases whose fp32 HR output exceeds this")
a = ap.parse_args()
dev = torch.device("cuda", 0)
peaks, src = bench.load_peaks()
peak = peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"]
hl = bench.gen_config()
Sup3rGan.seed(0)
model = Sup3rGan(hl, C.discriminator(3, "same", (2048, 1024)), precision=a.precision)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
rows = []
for chunk in [(8, 8, 12), (16, 16, 24), (32, 32, 48), (64, 64, 96)]:
    fl = bench.algorithmic_flops_per_chunk(hl, (*chunk, 4))
    for B in (1, 2, 4, 8, 16, 32):
        hr_gb = B * np.prod(chunk) * 25 * 12 * 4 * 4 / 1e9
        body_gb = B * np.prod(chunk) * 12 * 64 * 2 / 1e9
        if hr_gb > a.max_hr_gb:
            continue
        try:
            x = torch.randn((B, *chunk, 4), device=dev)
            plan = model.plan_for(model.generator, a.precision)
            plan.invalidate()
            for _ in range(2):
                plan.run_graphed(x)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); plan.run_graphed(x); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            vox = B * int(np.prod(chunk))
            tf = fl * B / (ms / 1e3) / 1e12
            rows.append((chunk, B, ms, vox / (ms / 1e3), tf, tf / peak, hr_gb))
            print(f"chunk {chunk} batch {B:2d}: {ms:8.2f} ms  {vox/(ms/1e3)/1e6:6.2f} M LR voxels/s  {tf:6.1f} TF/s  {100*tf/peak:5.1f} %")
        except Exception as e:
            print(f"chunk {chunk} batch {B}: failed: {repr(e)[:120]}")
        finally:
            plan.invalidate()
            torch.cuda.empty_cache()
if a.out:
    out = [f"# sweep (BASELINE configs[4]): north-star generator (5x/12x/4f), precision {a.precision}, one B200", "",
           f"Device-timed (CUDA events around one CUDA-graph replay, 256 MiB L2 flush before each, median of 5). "
           f"Algorithmic FLOPs per chunk as in bench.py; peak = {peak} TF/s ({src} sustained bf16).", "",
           "| LR chunk | batch | ms / step | M LR voxels/s | algorithmic TF/s | % of peak | fp32 HR output (GB) |",
           "|---|---|---|---|---|---|---|"]
    for chunk, B, ms, v, tf, fr, gb in rows:
        out.append(f"| {chunk[0]}x{chunk[1]}x{chunk[2]} | {B} | {ms:.2f} | {v/1e6:.2f} | {tf:.0f} | {100*fr:.1f} | {gb:.2f} |")
    open(a.out, "w").write("\n".join(out) + "\n")
