"""Run the narrow HR output conv (8 -> 4 ch, n x 80 x 80 x 288) a few times (for ncu): args n impl iters"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sup3r_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
impl = sys.argv[2] if len(sys.argv) > 2 else "mma"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
x = torch.randn((n, 80, 80, 288, 8), device=dev)
w = torch.randn((3, 3, 3, 8, 4), device=dev) * 0.1
b = torch.randn(4, device=dev)
spec = ops.ConvSpec(3, 8, 4, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1)
y = torch.empty((n, 80, 80, 288, 4), device=dev)
f = ops.conv_fwd_small_bf16 if impl == "mma" else ops.conv_fwd
for _ in range(iters):
    f(x, w, b, spec, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    f(x, w, b, spec, out=y)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / iters
vox = n * 80 * 80 * 288
print(f"{impl}: {us:.1f} us/launch, {vox * 48 / us / 1e3:.0f} GB/s algorithmic (48 B/voxel), {vox*2*216*4/us/1e6:.1f} TFLOP/s")
