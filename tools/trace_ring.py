"""Role-level clock64 breakdown of the ring kernel (CTA 0) for one 64->64 body conv launch:
  python tools/trace_ring.py [n] [precision bf16|fp16c] [flags]
flags = s3_umma_tuning.box_y experiment bits (8: no epilogue work, 16: generic MMA role)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sup3r_b200 import ops
from sup3r_b200._cabi import UmmaTuning
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16c"
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
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
spec = ops.ConvSpec(3, 64, 64, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1, act=2, alpha=0.2)
y = torch.empty((n, *dims, 64), device=dev)
y_hi = torch.empty_like(x_hi)
y_lo = torch.empty_like(x_hi)
res = torch.randn_like(y)
r_hi, r_lo = ops.pack_act_pad16(res, split=True, fmt=fmt)
cases = [("pad16", dict(want_f32=False, out_hi=y_hi, out_lo=y_lo if split else None)),
         ("res16_pair", dict(want_f32=False, out_hi=y_hi, out_lo=y_lo, res_hi=r_hi, res_lo=r_lo)),
         ("f32", dict(out=y))]
for name, kw in cases:
    trace = torch.zeros(64, dtype=torch.int64, device=dev)
    t = UmmaTuning(trace=trace.data_ptr(), box_y=flags)
    for _ in range(3):
        ops.conv_fwd_umma(x_hi, x_lo, w_hi, w_lo, b, spec, n, dims, tune=t, fmt=fmt, acc_scale=acc, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t2 = UmmaTuning(box_y=flags)
    e0.record()
    for _ in range(10):
        ops.conv_fwd_umma(x_hi, x_lo, w_hi, w_lo, b, spec, n, dims, tune=t2, fmt=fmt, acc_scale=acc, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    tr = trace.cpu().tolist()
    items = max(tr[5], 1)
    print(f"{prec} flags {flags} {name:12s}: {us:7.1f} us/launch ({2*n*16*16*288*27*64*64/us/1e6:6.1f} TF/s) "
          f"items/CTA {tr[5]} MMA role {tr[0]/items:.0f} cyc/item | epilogue warp: wait acc_full "
          f"{tr[8]/items:.0f}, work {tr[9]/items:.0f} per item")
