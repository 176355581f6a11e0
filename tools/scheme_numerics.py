"""CPU emulation of candidate tensor-core operand schemes on the north-star generator.

Every conv of the float64 torch restatement is replaced by sum_p conv(A_p, W_p) with the operands
rounded the way a scheme stores them (products / accumulation in float64, i.e. the emulation
isolates OPERAND rounding; the fp32 accumulator adds ~1e-6).  Reports max|d|/max|ref|,
rms(d)/rms(ref) and the 99.9th percentile of the element-wise relative error against the
un-rounded float64 run.  Needs no GPU:  python tools/scheme_numerics.py [s1 s2 t]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import torch_ref  # noqa: E402


def rnd(x, dt):
    return x.to(dt).to(torch.float64)


def f8(x, scale, dt=torch.float8_e4m3fn):
    lim = 448.0 if dt == torch.float8_e4m3fn else 57344.0
    return (x * scale).clamp(-lim, lim).to(torch.float32).to(dt).to(torch.float64) / scale


def p2(v):
    """largest power of two <= v"""
    return 2.0 ** np.floor(np.log2(v))


def split(x, w, scheme):
    """-> list of (A_p, W_p) operand pairs whose convs are summed."""
    h, b = torch.float16, torch.bfloat16
    if scheme == "exact":
        return [(x, w)]
    if scheme == "bf16":
        return [(rnd(x, b), rnd(w, b))]
    if scheme == "fp16":
        return [(rnd(x, h), rnd(w, h))]
    if scheme == "bf16x3":
        xh, wh = rnd(x, b), rnd(w, b)
        xl, wl = rnd(x - xh, b), rnd(w - wh, b)
        return [(xh, wh), (xl, wh), (xh, wl)]
    if scheme == "fp16_asplit":      # 2 passes: (A_hi + A_lo) * W_hi
        xh, wh = rnd(x, h), rnd(w, h)
        return [(xh, wh), (rnd(x - xh, h), wh)]
    if scheme == "fp16_wsplit":      # 2 passes: A_hi * (W_hi + W_lo)
        xh, wh = rnd(x, h), rnd(w, h)
        return [(xh, wh), (xh, rnd(w - wh, h))]
    if scheme == "bf16_asplit":
        xh, wh = rnd(x, b), rnd(w, b)
        return [(xh, wh), (rnd(x - xh, b), wh)]
    if scheme in ("fp16_f8corr", "fp16_f8corr_a", "fp16_f8corr_w", "fp16_f8corr_fixed"):
        # main fp16 pass + fp8 (e4m3) correction pass, K-concatenated: A_lo8*W8 + A8*W_lo8
        xh, wh = rnd(x, h), rnd(w, h)
        xl, wl = x - xh, w - wh
        if scheme.endswith("fixed"):
            sa_lo, sa = 2.0 ** 13, 1.0          # activations: static scales (|A| < 112 exact)
        else:
            sa_lo = p2(224.0 / max(float(xl.abs().max()), 1e-30))
            sa = p2(224.0 / max(float(x.abs().max()), 1e-30))
        sw = p2(224.0 / max(float(w.abs().max()), 1e-30))
        sw_lo = p2(224.0 / max(float(wl.abs().max()), 1e-30))
        out = [(xh, wh)]
        if not scheme.endswith("_w"):
            out.append((f8(xl, sa_lo), f8(w, sw)))
        if not scheme.endswith("_a"):
            out.append((f8(x, sa), f8(wl, sw_lo)))
        return out
    raise KeyError(scheme)


class SchemeNet(torch_ref.TorchRefNet):
    scheme = "exact"


def run(net, x, scheme):
    orig3, orig2 = F.conv3d, F.conv2d

    def mk(orig):
        def conv(xc, w, b=None, **kw):
            y = None
            for a_p, w_p in split(xc, w, scheme):
                t = orig(a_p, w_p, None, **kw)
                y = t if y is None else y + t
            if b is not None:
                y = y + b.reshape(1, -1, *([1] * (y.dim() - 2)))
            return y
        return conv
    torch_ref.F.conv3d, torch_ref.F.conv2d = mk(orig3), mk(orig2)
    try:
        return net(x).numpy()
    finally:
        torch_ref.F.conv3d, torch_ref.F.conv2d = orig3, orig2


def main():
    shp = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (8, 8, 6)
    from sup3r_b200.network import CustomNetwork
    CustomNetwork.seed(0)
    hl = bench.gen_config()
    net = CustomNetwork(hl, name="generator", device="cpu")
    net.build((1, *shp, 4))
    ws = [np.asarray(w.numpy(), np.float64) for w in net.weights]
    # biases are zero at init: give them values so that the bias path is live
    rng = np.random.default_rng(3)
    ws = [w if w.ndim > 1 else 0.05 * rng.standard_normal(w.shape) for w in ws]
    tnet = torch_ref.TorchRefNet(hl, ws, dtype=torch.float64)
    x = torch.as_tensor(np.random.default_rng(1).standard_normal((1, *shp, 4)))
    torch.set_num_threads(os.cpu_count())
    ref = run(tnet, x, "exact")
    print(f"LR {shp}: ref rms {np.sqrt(np.mean(ref ** 2)):.4f} max {np.abs(ref).max():.4f}")
    print(f"{'scheme':20s} {'passes':>6s} {'max/max':>10s} {'rms/rms':>10s} {'p99.9 elem-rel':>15s}")
    cost = {"bf16": 1, "fp16": 1, "bf16x3": 3, "fp16_asplit": 2, "fp16_wsplit": 2,
            "bf16_asplit": 2, "fp16_f8corr": 2, "fp16_f8corr_a": 1.5, "fp16_f8corr_w": 1.5,
            "fp16_f8corr_fixed": 2}
    for s, c in cost.items():
        y = run(tnet, x, s)
        d = np.abs(y - ref)
        er = np.sort((d / np.maximum(np.abs(ref), 1e-30)).ravel())
        print(f"{s:20s} {c:6.1f} {d.max() / np.abs(ref).max():10.2e} "
              f"{np.sqrt(np.mean(d ** 2)) / np.sqrt(np.mean(ref ** 2)):10.2e} "
              f"{er[int(0.999 * (er.size - 1))]:15.2e}")


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "fp16c"):
    main()


# ------------------------------------------------------------------------------------------
# Faithful emulation of the "fp16c" mode as the kernels execute it (fused plan steps): every
# tensor between tcgen05 layers is STORED as fp16 hi + e4m3 lo8 = (x - hi) * 2^A_SHIFT (+ an
# e4m3 copy a8 of x for the weight-residue term); weights are scaled by a per-layer power of
# two S so that max|W| S is in [2^13, 2^14): Wh = fp16(W S), W8 = e4m3(W S 2^-A_SHIFT),
# Wl8 = e4m3(W S - Wh);  acc = Ah Wh + Al8 W8 + A8 Wl8;  y = acc / S + bias.
def emulate_fp16c(hl, ws, x, a_shift=11, head="fp16", first="c", out="fp16", hr_store="fp16",
                  verbose=False):
    from sup3r_b200.network import CustomNetwork
    from sup3r_b200.plan import build_steps, FusedConv, SkipStep
    net = CustomNetwork(hl, name="generator", device="cpu")
    steps = build_steps(net.layers)
    e4 = torch.float8_e4m3fn

    def q8(t):
        return t.clamp(-448.0, 448.0).to(torch.float32).to(e4).to(torch.float64)

    def store(t):   # what the epilogue writes: (hi, lo8, a8); value seen by skips = hi + lo8 2^-a
        hi = t.clamp(-65504.0, 65504.0).to(torch.float16).to(torch.float64)
        lo8 = q8((t - hi) * 2.0 ** a_shift)
        return hi, lo8, q8(t)

    def conv(xc, w, mode):
        """xc channels-first reflect-padded operands tuple or tensor; mode 'c' / 'fp16' / 'exact'"""
        if mode == "exact":
            return F.conv3d(xc[3], w)
        hi, lo8, a8 = xc[:3]
        if mode == "fp16":
            return F.conv3d(hi, w.to(torch.float16).to(torch.float64))
        S = p2(2.0 ** 14 * 0.999 / float(w.abs().max()))
        wsc = w * S
        wh = wsc.to(torch.float16).to(torch.float64)
        w8 = q8(wsc * 2.0 ** -a_shift)
        wl8 = q8(wsc - wh)
        return (F.conv3d(hi, wh) + F.conv3d(lo8, w8) + F.conv3d(a8, wl8)) / S

    def run(mode_body, mode_edge):
        wi = iter(ws)
        cur = torch.as_tensor(x)
        skips = {}
        nconv = sum(isinstance(s, FusedConv) for s in steps)
        ci = 0
        for st in steps:
            if isinstance(st, FusedConv):
                k, b = next(wi), next(wi)
                k = torch.as_tensor(k)
                w = k.permute(4, 3, 0, 1, 2)
                cf = cur.permute(0, 4, 1, 2, 3)
                parts = store(cf) + (cf,)
                parts = tuple(F.pad(t, (1, 1, 1, 1, 1, 1), mode="reflect") for t in parts)
                body = k.shape[-2] == 64 and k.shape[-1] == 64
                mode = mode_body
                if mode_body != "exact" and not body:
                    mode = first if k.shape[-2] == 4 else (head if k.shape[-1] == 200 else out)
                y = conv(parts, w, mode) + torch.as_tensor(b).reshape(1, -1, 1, 1, 1)
                y = y.permute(0, 2, 3, 4, 1)
                if st.act:
                    y = F.leaky_relu(y, st.alpha)
                if st.m > 1:
                    y = torch.repeat_interleave(y, st.m, dim=3)
                if st.r > 1:
                    y = torch.stack([torch_ref.depth_to_space(y[:, :, :, i], st.r)
                                     for i in range(y.shape[3])], dim=3)
                if st.skip_add is not None:
                    y = y + skips.pop(st.skip_add)
                ci += 1
                if mode_body != "exact" and ci < nconv:
                    h, l8, _ = store(y)
                    if ci < nconv - 1:
                        y = h + l8 * 2.0 ** -a_shift
                    elif hr_store == "fp16":
                        y = h                      # HR tensor: fp16 only
                    elif hr_store == "pair":
                        y = h + (y - h).to(torch.float16).to(torch.float64)
                for name in st.skip_store:
                    skips[name] = y
                cur = y
            elif isinstance(st, SkipStep):
                if st.store:
                    skips[st.name] = cur
                else:
                    cur = cur + skips.pop(st.name)
            else:
                raise RuntimeError(f"unexpected eager step {st}")
        return cur.numpy()

    ref = run("exact", "exact")
    y = run("c", "fp16")
    d = np.abs(y - ref)
    er = np.sort((d / np.maximum(np.abs(ref), 1e-30)).ravel())
    return (d.max() / np.abs(ref).max(), np.sqrt(np.mean(d ** 2)) / np.sqrt(np.mean(ref ** 2)),
            er[int(0.999 * (er.size - 1))])


def main_fp16c():
    shp = (8, 8, 6)
    from sup3r_b200.network import CustomNetwork
    CustomNetwork.seed(0)
    hl = bench.gen_config()
    net = CustomNetwork(hl, name="generator", device="cpu")
    net.build((1, *shp, 4))
    ws = [np.asarray(w.numpy(), np.float64) for w in net.weights]
    rng = np.random.default_rng(3)
    ws = [w if w.ndim > 1 else 0.05 * rng.standard_normal(w.shape) for w in ws]
    x = np.random.default_rng(1).standard_normal((1, *shp, 4))
    torch.set_num_threads(os.cpu_count())
    for first, head, out, hr in (("fp16", "fp16", "fp16", "fp16"), ("c", "fp16", "fp16", "fp16"),
                                 ("c", "c", "fp16", "fp16"), ("c", "c", "exact", "fp16"),
                                 ("c", "c", "exact", "pair"), ("c", "fp16", "exact", "pair"),
                                 ("c", "c", "exact", "exact")):
        r = emulate_fp16c(hl, ws, x, a_shift=11, head=head, first=first, out=out, hr_store=hr)
        print(f"fp16c first {first:5s} head {head:5s} out {out:5s} hr {hr:5s}: max/max {r[0]:.2e} "
              f"rms/rms {r[1]:.2e} p99.9 elem-rel {r[2]:.2e}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "fp16c":
    main_fp16c()
