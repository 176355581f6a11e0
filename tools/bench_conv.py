"""Micro-benchmark of the convolution kernels at the north-star body shapes (CUDA events,
L2 flushed between iterations).  Writes gpurun_out/bench_conv.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sup3r_b200 import ops  # noqa: E402
from sup3r_b200._cabi import UmmaTuning  # noqa: E402


def time_fn(fn, iters=10, warm=3, flush=None, graph=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if graph:  # replay a captured launch so host-side launch latency is not in the window
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        fn = g.replay
        fn()
        torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    res = []
    shapes = [("body_b1", 1, (16, 16, 288), 64), ("body_b8", 8, (16, 16, 288), 64),
              ("head_b1", 1, (16, 16, 288), 200), ("body_20x20", 1, (20, 20, 1728), 64)]
    for name, n, dims, cout in shapes:
        x = torch.randn((n,) + dims + (64,), device=dev)
        w = torch.randn((3, 3, 3, 64, cout), device=dev) * 0.03
        b = torch.randn(cout, device=dev) * 0.1
        x_hi, x_lo = ops.pack_act_pad16(x, split=True)
        w_hi, w_lo = ops.pack_weights_umma(w, split=True)
        d2s = 5 if cout == 200 else 1
        spec = ops.ConvSpec(3, 64, cout, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1,
                            act=2, alpha=0.2, d2s=d2s)
        flops = 2.0 * n * dims[0] * dims[1] * dims[2] * 27 * 64 * cout
        _, od, oc = spec.out_dims(n, dims)
        y = torch.empty(ops._shape_from(n, od, oc, 3), device=dev)
        y_hi = torch.empty(ops.pad16_shape(n, od, oc, 3), device=dev, dtype=torch.bfloat16)
        y_lo = torch.empty_like(y_hi)
        res_t = torch.randn_like(y) if cout == 64 else None
        variants = [("f32out", dict(out=y, want_f32=True), False)]
        if cout == 64:
            variants += [("pad16out", dict(want_f32=False, out_hi=y_hi), False),
                         ("res_f32_pad16", dict(out=y, out_hi=y_hi, residual=res_t), False),
                         ("split_pad16", dict(want_f32=False, out_hi=y_hi, out_lo=y_lo), True)]
        for tiles in ([4, 2] if cout == 64 else [2, 1]):
            for ws in (3, 2):
                for vname, kw, split in variants:
                    t = UmmaTuning(tiles=tiles, w_stages=ws)

                    def fn():
                        ops.conv_fwd_umma(x_hi, x_lo if split else None, w_hi,
                                          w_lo if split else None, b, spec, n, dims, tune=t, **kw)
                    try:
                        med, best = time_fn(fn, flush=flush)
                    except Exception as e:  # noqa
                        res.append(dict(shape=name, variant=vname, tiles=tiles, ws=ws,
                                        error=repr(e)[:200]))
                        continue
                    r = dict(shape=name, variant=vname, tiles=tiles, ws=ws, ms=med, ms_best=best,
                             tflops=flops / med / 1e9, tflops_best=flops / best / 1e9)
                    print(r, flush=True)
                    res.append(r)
        if n == 1 and cout == 64:
            def fn2():
                ops.conv_fwd(x, w, b, spec, out=y)
            med, best = time_fn(fn2, iters=3, warm=1, flush=flush, graph=False)
            r = dict(shape=name, variant="direct_f32", ms=med, tflops=flops / med / 1e9)
            print(r, flush=True)
            res.append(r)
    # HBM-bound tail: 8 -> 4 channels at (80, 80, 288)
    x = torch.randn((1, 80, 80, 288, 8), device=dev)
    w = torch.randn((3, 3, 3, 8, 4), device=dev) * 0.1
    b = torch.randn(4, device=dev)
    spec = ops.ConvSpec(3, 8, 4, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1)
    y = torch.empty((1, 80, 80, 288, 4), device=dev)
    med, best = time_fn(lambda: ops.conv_fwd(x, w, b, spec, out=y), iters=5, warm=2, flush=flush)
    r = dict(shape="tail_8to4", variant="direct_f32", ms=med,
             gbs=(x.numel() + y.numel()) * 4 / med / 1e6)
    print(r, flush=True)
    res.append(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_conv.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
