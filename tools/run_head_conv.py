"""Time the 64 -> 200 head convolution (fused 5x depth_to_space) of the north-star generator:
  python tools/run_head_conv.py [n] [precision bf16|fp16c] [iters]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sup3r_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16c"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
fmt = 2 if prec == "fp16c" else 0
split = fmt == 2
dev = torch.device("cuda:0")
dims = (16, 16, 288)
x = torch.randn((n, *dims, 64), device=dev)
w = torch.randn((3, 3, 3, 64, 200), device=dev) * 0.03
b = torch.randn(200, device=dev) * 0.1
x_hi, x_lo = ops.pack_act_pad16(x, split=split, fmt=fmt)
pk = ops.pack_weights_umma(w, ndim=3, fmt=fmt)
w_hi, w_lo = pk[:2]
acc = pk[2] if len(pk) == 3 else 0.0
spec = ops.ConvSpec(3, 64, 200, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1, act=2,
                    alpha=0.2, d2s=5)
kw = dict(want_f32=False, want_map16=True) if prec == "bf16" else {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters + 2):
    if i == 2:
        e0.record()
    y = ops.conv_fwd_umma(x_hi, x_lo, w_hi, w_lo, b, spec, n, dims, fmt=fmt, acc_scale=acc, **kw)
e1.record()
torch.cuda.synchronize()
print(f"head {prec}: {e0.elapsed_time(e1) * 1e3 / iters:.1f} us/launch")
