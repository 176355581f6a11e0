"""Time the 64 -> 200 head conv with fused 5x depth_to_space (bf16 mapped output): args n flags iters"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sup3r_b200 import ops
from sup3r_b200._cabi import UmmaTuning
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
tiles = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dims = (16, 16, 288)
dev = torch.device("cuda:0")
x = torch.randn((n,) + dims + (64,), device=dev)
w = torch.randn((3, 3, 3, 64, 200), device=dev) * 0.03
b = torch.randn(200, device=dev) * 0.1
x_hi, _ = ops.pack_act_pad16(x)
w_hi, _ = ops.pack_weights_umma(w, ndim=3)
spec = ops.ConvSpec(3, 64, 200, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1, act=2, alpha=0.2, d2s=5)
t = UmmaTuning(ring_slots=flags, tiles=tiles)
y16 = torch.empty((n, 80, 80, 288, 8), device=dev, dtype=torch.bfloat16)
for _ in range(2):
    ops.conv_fwd_umma(x_hi, None, w_hi, None, b, spec, n, dims, want_f32=False, out_hi=y16, tune=t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.conv_fwd_umma(x_hi, None, w_hi, None, b, spec, n, dims, want_f32=False, out_hi=y16, tune=t)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / iters
print(f"head flags {flags} tiles {tiles}: {us:.1f} us/launch, {2*n*16*16*288*27*64*200/us/1e6:.1f} TF/s")
