"""Launch the 64->64 body convolution a few times (for ncu):
  python tools/run_body_conv.py [n] [precision bf16|fp16c] [plain|res] [iters] [flags]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sup3r_b200 import ops
from sup3r_b200._cabi import UmmaTuning
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16c"
variant = sys.argv[3] if len(sys.argv) > 3 else "plain"
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
fmt = 2 if prec == "fp16c" else 0
split = fmt == 2
dev = torch.device("cuda:0")
dims = (16, 16, 288)
x = torch.randn((n, *dims, 64), device=dev)
w = torch.randn((3, 3, 3, 64, 64), device=dev) * 0.03
b = torch.randn(64, device=dev) * 0.1
x_hi, x_lo = ops.pack_act_pad16(x, split=split, fmt=fmt)
pk = ops.pack_weights_umma(w, ndim=3, fmt=fmt)
w_hi, w_lo = pk[:2]
acc = pk[2] if len(pk) == 3 else 0.0
res = variant == "res"
spec = ops.ConvSpec(3, 64, 64, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1,
                    act=0 if res else 2, alpha=0.2)
y_hi = torch.empty_like(x_hi)
y_lo = torch.empty_like(x_hi)
r_hi, r_lo = ops.pack_act_pad16(torch.randn_like(x), split=True, fmt=fmt)
kw = dict(want_f32=False, out_hi=y_hi, out_lo=y_lo if (split or res) else None)
if res:
    kw.update(res_hi=r_hi, res_lo=r_lo)
t = UmmaTuning(box_y=flags)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters + 2):
    if i == 2:
        e0.record()
    ops.conv_fwd_umma(x_hi, x_lo, w_hi, w_lo, b, spec, n, dims, tune=t, fmt=fmt, acc_scale=acc, **kw)
e1.record()
torch.cuda.synchronize()
print(f"{prec} {variant} flags {flags}: {e0.elapsed_time(e1) * 1e3 / iters:.1f} us/launch")
