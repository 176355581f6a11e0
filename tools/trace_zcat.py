"""Role-level clock64 breakdown of the ring kernel (CTA 0): python tools/trace_zcat.py [n] [scheme] [flags]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sup3r_b200 import ops
from sup3r_b200._cabi import UmmaTuning
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scheme = int(sys.argv[2]) if len(sys.argv) > 2 else 0
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
dims = (16, 16, 288)
x = torch.randn((n, *dims, 64), device=dev)
w = torch.randn((3, 3, 3, 64, 64), device=dev) * 0.03
b = torch.randn(64, device=dev) * 0.1
if os.environ.get("S3_ZERO"):
    x.zero_(); w.zero_()
x_hi, _ = ops.pack_act_pad16(x)
w_hi, _ = ops.pack_weights_umma(w, ndim=3)
spec = ops.ConvSpec(3, 64, 64, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1, act=2, alpha=0.2)
y = torch.empty((n, *dims, 64), device=dev)
y_hi = torch.empty_like(x_hi)
y_lo = torch.empty_like(x_hi)
res = torch.randn_like(y)
r_hi, r_lo = ops.pack_act_pad16(res, split=True)
for name, kw in [("pad16", dict(want_f32=False, out_hi=y_hi)),
                 ("res16_pad16_lo", dict(want_f32=False, out_hi=y_hi, out_lo=y_lo, res_hi=r_hi, res_lo=r_lo)),
                 ("f32", dict(out=y)),
                 ("res_f32_pad16", dict(out=y, out_hi=y_hi, residual=res))]:
    trace = torch.zeros(64, dtype=torch.int64, device=dev)
    t = UmmaTuning(trace=trace.data_ptr(), scheme=scheme, box_y=flags)
    for _ in range(3):
        ops.conv_fwd_umma(x_hi, None, w_hi, None, b, spec, n, dims, tune=t, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t2 = UmmaTuning(scheme=scheme, box_y=flags)
    e0.record()
    for _ in range(10):
        ops.conv_fwd_umma(x_hi, None, w_hi, None, b, spec, n, dims, tune=t2, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    tr = trace.cpu().tolist()
    items = max(tr[5], 1)
    print(f"flags {flags} {name:16s}: {us:7.1f} us/launch ({2*n*16*16*288*27*64*64/us/1e6:6.1f} TF/s) items/CTA {tr[5]} "
          f"MMA warp {tr[0]/items:.0f} cyc/item | epilogue warp: wait acc_full {tr[8]/items:.0f}, work {tr[9]/items:.0f} (res-load waits {tr[10]/items:.0f}, store-read waits {tr[11]/items:.0f}; phases: tmem-ld {tr[12]/items:.0f}, +bias/act {tr[13]/items:.0f}, residual {tr[14]/items:.0f}, output {tr[15]/items:.0f}) per item")
