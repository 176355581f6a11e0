"""Launch the tcgen05 weight-gradient kernel a few times (timing / ncu target):
  python tools/run_wgrad.py [n] [z y x] [iters]     default: BASELINE configs[3] body shape"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sup3r_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dims = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (16, 16, 48)
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 10
dev = torch.device("cuda:0")
x = torch.randn((n, *dims, 64), device=dev)
dy = torch.randn((n, *dims, 64), device=dev)
x_hi, _ = ops.pack_act_pad16(x, split=False, fmt=ops.S3_FMT_FP16)
g_hi, _ = ops.pack_act_pad16(dy, split=False, fmt=ops.S3_FMT_FP16, halo=0)
spec = ops.ConvSpec(3, 64, 64, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1)
for _ in range(2):
    ops.conv_wgrad_umma(x_hi, g_hi, 1, n, dims, 64)
    ops.conv_wgrad(x, dy, spec, (3, 3, 3, 64, 64), want_bias=False)
torch.cuda.synchronize()
for name, fn in (("tcgen05", lambda: ops.conv_wgrad_umma(x_hi, g_hi, 1, n, dims, 64)),
                 ("fp32 CUDA-core", lambda: ops.conv_wgrad(x, dy, spec, (3, 3, 3, 64, 64), want_bias=False))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    fl = 2.0 * n * dims[0] * dims[1] * dims[2] * 27 * 64 * 64
    print(f"wgrad {n} x {dims} 64->64 {name}: {us:.1f} us/launch, {fl / us / 1e6:.1f} algorithmic TFLOP/s")
