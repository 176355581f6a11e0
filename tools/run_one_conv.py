"""Run one tcgen05 conv configuration a few times (for ncu): args n cout tiles ws variant iters"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sup3r_b200 import ops  # noqa: E402
from sup3r_b200._cabi import UmmaTuning  # noqa: E402

n, cout, tiles, ws = (int(a) for a in sys.argv[1:5])
variant = sys.argv[5]
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
dims = (16, 16, 288)
dev = torch.device("cuda:0")
x = torch.randn((n,) + dims + (64,), device=dev)
w = torch.randn((3, 3, 3, 64, cout), device=dev) * 0.03
b = torch.randn(cout, device=dev) * 0.1
split = variant.startswith("split")
x_hi, x_lo = ops.pack_act_pad16(x, split=split)
w_hi, w_lo = ops.pack_weights_umma(w, split=split)
spec = ops.ConvSpec(3, 64, cout, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1, act=2,
                    alpha=0.2, d2s=5 if cout == 200 else 1)
_, od, oc = spec.out_dims(n, dims)
y = torch.empty(ops._shape_from(n, od, oc, 3), device=dev)
y_hi = torch.empty(ops.pad16_shape(n, od, oc, 3), device=dev, dtype=torch.bfloat16)
y_lo = torch.empty_like(y_hi) if split else None
kw = dict(out=y) if variant == "f32out" else dict(want_f32=False, out_hi=y_hi, out_lo=y_lo)
scheme = int(sys.argv[7]) if len(sys.argv) > 7 else 0
flags = int(sys.argv[8]) if len(sys.argv) > 8 else 0
t = UmmaTuning(tiles=tiles, w_stages=ws, scheme=scheme, box_y=flags)
for _ in range(2):
    ops.conv_fwd_umma(x_hi, x_lo, w_hi, w_lo, b, spec, n, dims, tune=t, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.conv_fwd_umma(x_hi, x_lo, w_hi, w_lo, b, spec, n, dims, tune=t, **kw)
e1.record()
torch.cuda.synchronize()
print(f"{variant} tiles={tiles} ws={ws}: {e0.elapsed_time(e1) * 1e3 / iters:.1f} us/launch")
