"""Parity of every C-ABI kernel against the CPU oracle (numpy restatement of the reference's
layer semantics, oracle/layers_ref.py) on seeded inputs.  fp32 kernels: rtol 1e-4 of the tensor
scale (the north-star bound is 1e-3 relative)."""
import zlib

import numpy as np
import pytest
import torch

from oracle import layers_ref as L

pytestmark = pytest.mark.gpu

TOL = 1e-4


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def dev(a, cuda):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=cuda)


def rng_arr(rng, shape, scale=1.0):
    return (rng.standard_normal(shape) * scale).astype(np.float32)


CONV_CASES = [
    # ndim, in shape, cout, k, stride, padding, act
    (3, (1, 6, 7, 9, 4), 64, 3, 1, "valid", None),
    (3, (2, 8, 8, 8, 64), 64, 3, 1, "same", "leaky"),
    (3, (1, 9, 10, 11, 6), 32, 3, 2, "same", "leaky"),
    (3, (1, 9, 10, 11, 32), 20, 3, 2, "valid", None),
    (2, (3, 12, 13, 2), 64, 3, 1, "valid", "relu"),
    (2, (2, 11, 12, 64), 256, 3, 2, "same", "leaky"),
    (2, (2, 20, 20, 8), 4, 3, 1, "same", None),
    (3, (1, 5, 5, 12, 18), 14, 3, 1, "same", None),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_matches_oracle(cuda, case):
    from sup3r_b200 import ops
    ndim, shape, cout, k, s, padding, act = case
    rng = np.random.default_rng(1)
    x = rng_arr(rng, shape)
    w = rng_arr(rng, (k,) * ndim + (shape[-1], cout), 0.1)
    b = rng_arr(rng, (cout,), 0.1)
    ref = L.conv_nd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), s, padding)
    if act == "leaky":
        ref = np.where(ref >= 0, ref, 0.2 * ref)
    elif act == "relu":
        ref = np.maximum(ref, 0)
    sp = shape[1:-1]
    if padding == "same":
        pads = [L.same_pads(n, k, s) for n in sp]
    else:
        pads = [(0, 0)] * ndim
    z = [(0, 0)] * (3 - ndim)
    spec = ops.ConvSpec(ndim, shape[-1], cout, (1,) * (3 - ndim) + (k,) * ndim,
                        stride=(1,) * (3 - ndim) + (s,) * ndim,
                        pad_lo=tuple(p[0] for p in z + pads), pad_hi=tuple(p[1] for p in z + pads),
                        pad_mode=0, act={None: 0, "relu": 1, "leaky": 2}[act], alpha=0.2)
    y = ops.conv_fwd(dev(x, cuda), dev(w, cuda), dev(b, cuda), spec).cpu().numpy()
    assert y.shape == ref.shape
    assert rel_err(y, ref) < TOL


@pytest.mark.parametrize("ndim,shape,cout", [(3, (1, 5, 6, 7, 3), 8), (2, (2, 9, 8, 5), 12),
                                             (3, (1, 4, 4, 6, 64), 64)])
def test_fused_reflect_conv_equals_pad_conv_crop(cuda, ndim, shape, cout):
    """FlexiblePadding(3, REFLECT) -> Conv(valid) -> Cropping(2) == reflect-1 fused conv."""
    from sup3r_b200 import ops
    rng = np.random.default_rng(7)
    x = rng_arr(rng, shape)
    w = rng_arr(rng, (3,) * ndim + (shape[-1], cout), 0.1)
    b = rng_arr(rng, (cout,), 0.1)
    pads = [[0, 0]] + [[3, 3]] * ndim + [[0, 0]]
    ref = L.crop_nd(L.conv_nd(L.tf_pad(x.astype(np.float64), pads, "REFLECT"),
                              w.astype(np.float64), b.astype(np.float64)), 2)
    one = (0,) * (3 - ndim) + (1,) * ndim
    spec = ops.ConvSpec(ndim, shape[-1], cout, (1,) * (3 - ndim) + (3,) * ndim, pad_lo=one,
                        pad_hi=one, pad_mode=1)
    y = ops.conv_fwd(dev(x, cuda), dev(w, cuda), dev(b, cuda), spec).cpu().numpy()
    assert rel_err(y, ref) < TOL


@pytest.mark.parametrize("cin,cout,shape,mode", [(8, 4, (1, 9, 13, 70), 1), (4, 2, (2, 4, 8, 32), 1),
                                                 (12, 6, (1, 5, 9, 33), 0), (3, 1, (1, 6, 6, 6), 1),
                                                 (32, 2, (1, 4, 8, 40), 1), (18, 14, (1, 4, 5, 6), 1)])
def test_small_channel_conv_matches_oracle(cuda, cin, cout, shape, mode):
    """The narrow-layer kernel (conv_small.cu) incl. tile-edge masking, residual and affine;
    (18 -> 14 falls back to the generic kernel)."""
    from sup3r_b200 import ops
    rng = np.random.default_rng(21)
    x = rng_arr(rng, shape + (cin,))
    w = rng_arr(rng, (3, 3, 3, cin, cout), 0.1)
    b = rng_arr(rng, (cout,), 0.1)
    res = rng_arr(rng, shape + (cout,))
    sc, sh = rng_arr(rng, (cout,)), rng_arr(rng, (cout,))
    pads = [[0, 0]] + [[1, 1]] * 3 + [[0, 0]]
    xp = L.tf_pad(x.astype(np.float64), pads, "REFLECT" if mode == 1 else "CONSTANT")
    ref = L.conv_nd(xp, w.astype(np.float64), b.astype(np.float64))
    ref = (np.where(ref >= 0, ref, 0.2 * ref) + res) * sc + sh
    spec = ops.ConvSpec(3, cin, cout, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=mode,
                        act=2, alpha=0.2)
    y = ops.conv_fwd(dev(x, cuda), dev(w, cuda), dev(b, cuda), spec, residual=dev(res, cuda),
                     post_scale=dev(sc, cuda), post_shift=dev(sh, cuda)).cpu().numpy()
    assert y.shape == ref.shape and rel_err(y, ref) < TOL


def test_conv_transpose_as_flipped_conv(cuda):
    """Conv2DTranspose(valid, stride 1) == zero-pad-2 conv with the flipped, swapped kernel."""
    from sup3r_b200 import ops
    rng = np.random.default_rng(3)
    x = rng_arr(rng, (2, 7, 8, 5))
    wt = rng_arr(rng, (3, 3, 6, 5), 0.1)  # (kh, kw, cout, cin)
    b = rng_arr(rng, (6,), 0.1)
    ref = L.conv_transpose_nd(x.astype(np.float64), wt.astype(np.float64), b.astype(np.float64))
    w = np.ascontiguousarray(wt[::-1, ::-1].transpose(0, 1, 3, 2))
    spec = ops.ConvSpec(2, 5, 6, (1, 3, 3), pad_lo=(0, 2, 2), pad_hi=(0, 2, 2), pad_mode=0)
    y = ops.conv_fwd(dev(x, cuda), dev(w, cuda), dev(b, cuda), spec).cpu().numpy()
    assert y.shape == ref.shape and rel_err(y, ref) < TOL


@pytest.mark.parametrize("r,m,method,roll,rep", [(2, 1, 0, 0, (1, 1, 1)), (1, 3, 1, 2, (1, 1, 1)),
                                                 (1, 1, 0, 0, (1, 1, 3)), (5, 1, 0, 0, (1, 1, 1)),
                                                 (2, 2, 1, -1, (1, 1, 1))])
def test_conv_epilogue_scatter_3d(cuda, r, m, method, roll, rep):
    """conv -> LeakyReLU -> SpatioTemporalExpansion fused == oracle layer sequence; also the
    16-bit padded + mirrored copy equals a REFLECT pad of the result."""
    from sup3r_b200 import ops
    rng = np.random.default_rng(11)
    cmap = 3
    cout = cmap * r * r * m
    x = rng_arr(rng, (2, 4, 5, 6, 7))
    w = rng_arr(rng, (3, 3, 3, 7, cout), 0.1)
    b = rng_arr(rng, (cout,), 0.1)
    ref = L.conv_nd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), 1, "same")
    ref = np.where(ref >= 0, ref, 0.2 * ref)
    if rep[2] > 1:
        ref = L.spatiotemporal_expansion(ref, 1, rep[2], "nearest")
    else:
        ref = L.spatiotemporal_expansion(ref, r, m, "depth_to_time", roll)
    spec = ops.ConvSpec(3, 7, cout, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=0,
                        act=2, alpha=0.2, d2s=r, d2t=m, t_roll=roll, out_repeat=rep)
    y, y_hi, y_lo = ops.conv_fwd(dev(x, cuda), dev(w, cuda), dev(b, cuda), spec, want_pad16=True,
                                 split=True)
    assert tuple(y.shape) == ref.shape
    assert rel_err(y.cpu().numpy(), ref) < TOL
    full = (y_hi.float() + y_lo.float()).cpu().numpy()
    ref_pad = np.pad(ref, [(0, 0), (1, 1), (1, 1), (1, 1), (0, 0)], mode="reflect")
    assert rel_err(full, ref_pad) < 2e-5  # two bf16 terms carry ~16 mantissa bits


def test_conv_epilogue_d2s_2d_and_concat_stride(cuda):
    from sup3r_b200 import ops
    rng = np.random.default_rng(12)
    x = rng_arr(rng, (2, 6, 7, 5))
    w = rng_arr(rng, (3, 3, 5, 16), 0.1)
    b = rng_arr(rng, (16,), 0.1)
    ref = L.depth_to_space(L.conv_nd(x.astype(np.float64), w.astype(np.float64),
                                     b.astype(np.float64), 1, "same"), 2)
    spec = ops.ConvSpec(2, 5, 16, (1, 3, 3), pad_lo=(0, 1, 1), pad_hi=(0, 1, 1), d2s=2,
                        out_cstride=6, out_coffset=1)
    out = torch.full((2, 12, 14, 6), -7.0, device=cuda)
    ops.conv_fwd(dev(x, cuda), dev(w, cuda), dev(b, cuda), spec, out=out)
    o = out.cpu().numpy()
    assert rel_err(o[..., 1:5], ref) < TOL
    assert np.all(o[..., 0] == -7.0) and np.all(o[..., 5] == -7.0)


def test_conv_residual_and_affine(cuda):
    from sup3r_b200 import ops
    rng = np.random.default_rng(13)
    x = rng_arr(rng, (1, 4, 5, 6, 8))
    w = rng_arr(rng, (3, 3, 3, 8, 8), 0.1)
    b = rng_arr(rng, (8,), 0.1)
    res = rng_arr(rng, (1, 4, 5, 6, 8))
    sc, sh = rng_arr(rng, (8,)), rng_arr(rng, (8,))
    ref = (L.conv_nd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), 1, "same")
           + res) * sc + sh
    spec = ops.ConvSpec(3, 8, 8, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1))
    y = ops.conv_fwd(dev(x, cuda), dev(w, cuda), dev(b, cuda), spec, residual=dev(res, cuda),
                     post_scale=dev(sc, cuda), post_shift=dev(sh, cuda)).cpu().numpy()
    assert rel_err(y, ref) < TOL


@pytest.mark.parametrize("case", [(3, (2, 6, 7, 8, 5), 9, 1, "same"), (3, (1, 9, 8, 7, 6), 32, 2, "same"),
                                  (2, (2, 11, 10, 64), 70, 2, "valid"), (2, (1, 9, 9, 3), 4, 1, "valid")])
def test_conv_backward_matches_autograd(cuda, case):
    """dgrad / wgrad / dbias against float64 torch autograd of the same cross-correlation."""
    from sup3r_b200 import ops
    import torch.nn.functional as F
    ndim, shape, cout, s, padding = case
    rng = np.random.default_rng(5)
    x = rng_arr(rng, shape)
    w = rng_arr(rng, (3,) * ndim + (shape[-1], cout), 0.1)
    sp = shape[1:-1]
    pads = [L.same_pads(n, 3, s) for n in sp] if padding == "same" else [(0, 0)] * ndim
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    wt = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    perm_in = (0, ndim + 1) + tuple(range(1, ndim + 1))
    xin = xt.permute(*perm_in)
    fpad = []
    for lo, hi in reversed(pads):
        fpad += [lo, hi]
    xin = F.pad(xin, fpad)
    wk = wt.permute(ndim + 1, ndim, *range(ndim))
    conv = F.conv3d if ndim == 3 else F.conv2d
    yt = conv(xin, wk, stride=s)
    yt = yt.permute(0, *range(2, ndim + 2), 1)
    dy = rng_arr(rng, tuple(yt.shape))
    yt.backward(torch.tensor(dy, dtype=torch.float64))
    z = [(0, 0)] * (3 - ndim)
    spec = ops.ConvSpec(ndim, shape[-1], cout, (1,) * (3 - ndim) + (3,) * ndim,
                        stride=(1,) * (3 - ndim) + (s,) * ndim,
                        pad_lo=tuple(p[0] for p in z + pads), pad_hi=tuple(p[1] for p in z + pads))
    dx = ops.conv_dgrad(dev(dy, cuda), dev(w, cuda), spec, shape).cpu().numpy()
    dw, db = ops.conv_wgrad(dev(x, cuda), dev(dy, cuda), spec, w.shape)
    assert rel_err(dx, xt.grad.numpy()) < TOL
    assert rel_err(dw.cpu().numpy(), wt.grad.numpy()) < TOL
    assert rel_err(db.cpu().numpy(), dy.reshape(-1, cout).sum(0)) < TOL


@pytest.mark.parametrize("mode", ["REFLECT", "CONSTANT", "SYMMETRIC"])
def test_pad_fwd_bwd(cuda, mode):
    from sup3r_b200 import ops
    rng = np.random.default_rng(2)
    x = rng_arr(rng, (2, 5, 6, 7, 3))
    pads = [[0, 0], [3, 3], [2, 1], [3, 0], [0, 0]]
    ref = L.tf_pad(x, pads, mode)
    code = ops.PAD_CODES[mode]
    y = ops.pad_fwd(dev(x, cuda), pads, code)
    assert np.array_equal(y.cpu().numpy(), ref)
    # adjoint identity <pad(x), dy> == <x, pad_bwd(dy)>
    dy = rng_arr(rng, ref.shape)
    dx = ops.pad_bwd(dev(dy, cuda), x.shape, pads, code).cpu().numpy()
    lhs = float((ref.astype(np.float64) * dy).sum())
    rhs = float((x.astype(np.float64) * dx).sum())
    assert abs(lhs - rhs) < 1e-4 * max(abs(lhs), 1.0)
    # 4 | channels takes the float4 kernel: same sums as the scalar kernel, channel by channel
    dy8 = rng_arr(rng, ref.shape[:-1] + (8,))
    got = ops.pad_bwd(dev(dy8, cuda), x.shape[:-1] + (8,), pads, code)
    for ch in (0, 3, 7):        # (a 1-channel tensor goes through the scalar kernel)
        one = ops.pad_bwd(dev(np.ascontiguousarray(dy8[..., ch:ch + 1]), cuda),
                          x.shape[:-1] + (1,), pads, code)
        assert torch.equal(got[..., ch:ch + 1], one)


def test_crop_act_add_concat_affine(cuda):
    from sup3r_b200 import ops
    rng = np.random.default_rng(4)
    x = rng_arr(rng, (2, 6, 7, 8, 5))
    y = ops.crop_fwd(dev(x, cuda), [(0, 0), (2, 2), (1, 2), (0, 3), (0, 0)]).cpu().numpy()
    assert np.array_equal(y, x[:, 2:-2, 1:-2, :-3])
    dx = ops.crop_bwd(dev(y, cuda), x.shape, [(0, 0), (2, 2), (1, 2), (0, 3), (0, 0)]).cpu().numpy()
    ref = np.zeros_like(x)
    ref[:, 2:-2, 1:-2, :-3] = y
    assert np.array_equal(dx, ref)
    for act, name in [(1, "relu"), (2, "leaky_relu"), (3, "sigmoid"), (4, "tanh")]:
        a = ops.act_fwd(dev(x, cuda), act, 0.2).cpu().numpy()
        assert rel_err(a, L.activation(x.astype(np.float64), name, 0.2)) < 1e-5
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    out = torch.nn.functional.leaky_relu(xt, 0.2)
    g = rng_arr(rng, x.shape)
    out.backward(torch.tensor(g, dtype=torch.float64))
    got = ops.act_bwd(dev(out.detach().numpy(), cuda), dev(g, cuda), 2, 0.2).cpu().numpy()
    assert rel_err(got, xt.grad.numpy()) < 1e-6
    b = rng_arr(rng, x.shape)
    assert np.array_equal(ops.add(dev(x, cuda), dev(b, cuda)).cpu().numpy(), x + b)
    e = rng_arr(rng, x.shape[:-1] + (1,))
    cat = ops.concat_fwd(dev(x, cuda), dev(e, cuda)).cpu().numpy()
    assert np.array_equal(cat, np.concatenate((x, e), -1))
    da, db = ops.concat_bwd(dev(cat, cuda), 5, 1, want_b=True)
    assert np.array_equal(da.cpu().numpy(), x) and np.array_equal(db.cpu().numpy(), e)
    sc, sh = rng_arr(rng, (5,)), rng_arr(rng, (5,))
    aff = ops.channel_affine(dev(x, cuda), dev(sc, cuda), dev(sh, cuda)).cpu().numpy()
    assert rel_err(aff, x * sc + sh) < 1e-6


@pytest.mark.parametrize("r,m,method,roll", [(2, 1, 0, 0), (1, 3, 0, 0), (3, 2, 0, 0),
                                             (1, 4, 1, 2), (2, 2, 1, -3), (5, 1, 0, 0)])
def test_expand_3d_fwd_bwd(cuda, r, m, method, roll):
    from sup3r_b200 import ops
    rng = np.random.default_rng(6)
    c = 2 * r * r * (m if method == 1 else 1)
    x = rng_arr(rng, (2, 3, 4, 5, c))
    meth = "depth_to_time" if method == 1 else "nearest"
    ref = L.spatiotemporal_expansion(x, r, m, meth, roll)
    y = ops.expand_fwd(dev(x, cuda), r, m, method, roll)
    assert np.array_equal(y.cpu().numpy(), ref)
    dy = rng_arr(rng, ref.shape)
    dx = ops.expand_bwd(dev(dy, cuda), x.shape, r, m, method, roll).cpu().numpy()
    lhs = float((ref.astype(np.float64) * dy).sum())
    rhs = float((x.astype(np.float64) * dx).sum())
    assert abs(lhs - rhs) < 1e-4 * max(abs(lhs), 1.0)


def test_expand_2d(cuda):
    from sup3r_b200 import ops
    rng = np.random.default_rng(8)
    x = rng_arr(rng, (3, 4, 5, 18))
    ref = L.depth_to_space(x, 3)
    y = ops.expand_fwd(dev(x, cuda), 3)
    assert np.array_equal(y.cpu().numpy(), ref)
    dx = ops.expand_bwd(dev(ref, cuda), x.shape, 3).cpu().numpy()
    assert np.array_equal(dx, x)


def test_dense_fwd_bwd(cuda):
    from sup3r_b200 import ops
    rng = np.random.default_rng(9)
    m, k, n = 5, 777, 130
    x, w, b = rng_arr(rng, (m, k)), rng_arr(rng, (k, n), 0.05), rng_arr(rng, (n,))
    ref = x.astype(np.float64) @ w + b
    y = ops.dense_fwd(dev(x, cuda), dev(w, cuda), dev(b, cuda), 2, 0.2).cpu().numpy()
    assert rel_err(y, np.where(ref >= 0, ref, 0.2 * ref)) < TOL
    dy = rng_arr(rng, (m, n))
    dx, dw, db = ops.dense_bwd(dev(x, cuda), dev(w, cuda), dev(dy, cuda))
    assert rel_err(dx.cpu().numpy(), dy.astype(np.float64) @ w.T) < TOL
    assert rel_err(dw.cpu().numpy(), x.T.astype(np.float64) @ dy) < TOL
    assert rel_err(db.cpu().numpy(), dy.sum(0)) < TOL


def test_losses_and_adam(cuda):
    from sup3r_b200 import ops
    rng = np.random.default_rng(10)
    g, t = rng_arr(rng, (2, 6, 6, 4, 5)), rng_arr(rng, (2, 6, 6, 4, 5))
    for kind in (0, 1):
        gt = torch.tensor(g, dtype=torch.float64, requires_grad=True)
        d = gt[..., :4] - torch.tensor(t, dtype=torch.float64)[..., :4]
        ref = (d * d).mean() if kind == 0 else d.abs().mean()
        (0.7 * ref).backward()
        loss, dg = ops.content_loss(dev(g, cuda), dev(t, cuda), 4, kind, 0.7, want_grad=True)
        assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
        assert rel_err(dg.cpu().numpy(), gt.grad.numpy()) < 1e-5
    a, b = rng_arr(rng, (9, 1)), rng_arr(rng, (9, 1))
    at = torch.tensor(a, dtype=torch.float64, requires_grad=True)
    bt = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    logits = torch.cat([at - bt.mean(), bt - at.mean()], 0)
    labels = torch.cat([torch.ones_like(at), torch.zeros_like(bt)], 0)
    ref = torch.nn.functional.binary_cross_entropy_with_logits(logits, labels)
    (0.3 * ref).backward()
    loss, da, db = ops.loss_disc(dev(a, cuda), dev(b, cuda), 0.3, want_grad=True)
    assert abs(loss.item() - ref.item()) < 1e-5
    assert rel_err(da.cpu().numpy(), at.grad.numpy()) < 1e-4
    assert rel_err(db.cpu().numpy(), bt.grad.numpy()) < 1e-4
    # keras Adam, two steps
    p0, gr = rng_arr(rng, (1000,)), rng_arr(rng, (1000,))
    p = dev(p0, cuda)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    pr, mr, vr = p0.astype(np.float64), np.zeros(1000), np.zeros(1000)
    for step in (1, 2):
        ops.adam_step(p, dev(gr, cuda), m, v, 1e-3, 0.9, 0.999, 1e-7, step)
        mr = 0.9 * mr + 0.1 * gr
        vr = 0.999 * vr + 0.001 * gr.astype(np.float64) ** 2
        lr_t = 1e-3 * np.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step)
        pr = pr - lr_t * mr / (np.sqrt(vr) + 1e-7)
    assert rel_err(p.cpu().numpy(), pr) < 1e-5


def test_stats_and_channel_check(cuda):
    from sup3r_b200 import ops
    rng = np.random.default_rng(14)
    x = rng_arr(rng, (3, 5, 5, 4))
    x[..., 2] = 1.5
    x[0, 0, 0, 3] = np.nan
    st = ops.stats(dev(x, cuda)).cpu().numpy()
    fin = x[np.isfinite(x)]
    assert abs(st[0] - fin.sum()) < 1e-3 and st[2] == 1
    assert st[3] == fin.min() and st[4] == fin.max()
    cc = ops.channel_check(dev(x, cuda)).cpu().numpy()
    assert cc[2, 0] == cc[2, 1] == 1.5 and cc[3, 2] == 1 and cc[0, 2] == 0
    assert cc[0, 0] == x[..., 0].min() and cc[1, 1] == x[..., 1].max()


# ----------------------------------------------------------------------------- tcgen05 conv
def _bf16_exact(a):
    return torch.as_tensor(a).to(torch.bfloat16).to(torch.float32).numpy()


UMMA_CASES = [
    # n, (z, y, x), variant
    (1, (16, 16, 24), "pad16"),          # hot shape: ring kernel, straight-line MMA role, v2 epilogue
    (2, (8, 16, 40), "pad16_lo"),
    (1, (16, 16, 24), "res16"),
    (2, (4, 13, 21), "res16"),           # ragged y / x tiles
    (1, (12, 20, 9), "res16_nolo"),
    (1, (6, 9, 17), "pad16"),            # planes % 4 != 0 -> generic MMA role
    (1, (16, 16, 24), "f32"),            # fp32 destination (thread-per-row epilogue)
    (1, (8, 16, 24), "res_f32"),
]


@pytest.mark.parametrize("n,dims,variant", UMMA_CASES)
def test_umma_conv_matches_direct(cuda, n, dims, variant):
    """tcgen05 64->64 3x3x3 reflect conv (+bias, LeakyReLU, SkipConnection add) against the fp32
    direct kernel on bf16-exact operands: products are exact, only the fp32 summation order
    differs.  16-bit outputs are compared INCLUDING the REFLECT halo the epilogue writes."""
    from sup3r_b200 import ops
    rng = np.random.default_rng(zlib.crc32(repr((n, dims, variant)).encode()))
    x = _bf16_exact(rng_arr(rng, (n, *dims, 64)))
    w = _bf16_exact(rng_arr(rng, (3, 3, 3, 64, 64), 0.05))
    b = rng_arr(rng, (64,), 0.1)
    res = rng_arr(rng, (n, *dims, 64))
    xd, wd, bd, rd = dev(x, cuda), dev(w, cuda), dev(b, cuda), dev(res, cuda)
    has_res = variant.startswith("res")
    spec = ops.ConvSpec(3, 64, 64, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1,
                        act=0 if has_res else 2, alpha=0.2)
    x_hi, _ = ops.pack_act_pad16(xd)
    w_hi, _ = ops.pack_weights_umma(wd, ndim=3)
    r_hi, r_lo = ops.pack_act_pad16(rd, split=True)
    if variant == "res16_nolo":
        r_lo = None
    if has_res:   # the addend the kernel sees
        rd_eff = ops.unpack_act_pad16(r_hi, r_lo, 3) if variant != "res_f32" else rd
    ref = ops.conv_fwd(xd, wd, bd, spec, residual=rd_eff if has_res else None)
    if variant in ("f32", "res_f32"):
        y, _, _ = ops.conv_fwd_umma(x_hi, None, w_hi, None, bd, spec, n, dims,
                                    residual=rd if has_res else None)
        assert rel_err(y.cpu().numpy(), ref.cpu().numpy()) < 2e-5
        return
    want_lo = variant in ("pad16_lo", "res16")
    _, y_hi, y_lo = ops.conv_fwd_umma(x_hi, None, w_hi, None, bd, spec, n, dims, want_f32=False,
                                      want_pad16=True, want_lo=want_lo,
                                      res_hi=r_hi if has_res else None,
                                      res_lo=r_lo if has_res else None)
    ref_hi, ref_lo = ops.pack_act_pad16(ref, split=True)
    scale = float(ref.abs().max())
    # hi: bf16 rounding of an fp32 sum that may differ in the last bits -> allow one bf16 ulp
    dh = (y_hi.float() - ref_hi.float()).abs().max().item()
    assert dh <= scale * 2.0 ** -7, dh
    frac_equal = (y_hi == ref_hi).float().mean().item()
    assert frac_equal > 0.995, frac_equal
    if want_lo:
        got = ops.unpack_act_pad16(y_hi, y_lo, 3)
        assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < 5e-5
        full = y_hi.float() + y_lo.float()
        full_ref = ref_hi.float() + ref_lo.float()
        assert (full - full_ref).abs().max().item() <= scale * 1e-4   # halos of hi + lo too


@pytest.mark.parametrize("cin,cout,shape,mode", [(8, 4, (1, 9, 21, 70), 1), (4, 2, (2, 4, 8, 32), 1),
                                                   (8, 8, (1, 5, 17, 33), 0), (6, 3, (1, 6, 6, 40), 2)])
def test_small_channel_mma_conv(cuda, cin, cout, shape, mode):
    """Narrow 3x3x3 conv on mma.sync (bf16 operands) vs the fp32 direct kernel on bf16-exact
    inputs: identical products, fp32 summation order differs."""
    from sup3r_b200 import ops
    rng = np.random.default_rng(cin * 100 + cout)
    x = _bf16_exact(rng_arr(rng, (*shape, cin)))
    w = _bf16_exact(rng_arr(rng, (3, 3, 3, cin, cout), 0.2))
    b = rng_arr(rng, (cout,), 0.1)
    sc, sh = rng_arr(rng, (cout,)), rng_arr(rng, (cout,))
    spec = ops.ConvSpec(3, cin, cout, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=mode)
    ref = ops.conv_fwd(dev(x, cuda), dev(w, cuda), dev(b, cuda), spec, post_scale=dev(sc, cuda),
                       post_shift=dev(sh, cuda))
    assert ops.small_bf16_ok(spec)
    y = ops.conv_fwd_small_bf16(dev(x, cuda), dev(w, cuda), dev(b, cuda), spec,
                                post_scale=dev(sc, cuda), post_shift=dev(sh, cuda))
    assert rel_err(y.cpu().numpy(), ref.cpu().numpy()) < 2e-5


def test_umma_d2s_16bit_destination_feeds_small_conv(cuda):
    """64 -> 200 head with fused 5x depth_to_space writing an unpadded bf16 tensor, consumed by the
    narrow tensor-core conv (bf16 input path): both against the fp32 kernels."""
    from sup3r_b200 import ops
    rng = np.random.default_rng(5)
    n, dims = 1, (4, 5, 24)
    x = _bf16_exact(rng_arr(rng, (n, *dims, 64)))
    w = _bf16_exact(rng_arr(rng, (3, 3, 3, 64, 200), 0.05))
    b = rng_arr(rng, (200,), 0.1)
    xd, wd, bd = dev(x, cuda), dev(w, cuda), dev(b, cuda)
    spec = ops.ConvSpec(3, 64, 200, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1,
                        d2s=5)
    ref = ops.conv_fwd(xd, wd, bd, spec)                      # (1, 20, 25, 24, 8) f32
    x_hi, _ = ops.pack_act_pad16(xd)
    w_hi, _ = ops.pack_weights_umma(wd, ndim=3)
    _, y16, _ = ops.conv_fwd_umma(x_hi, None, w_hi, None, bd, spec, n, dims, want_f32=False,
                                  want_map16=True)
    assert y16.dtype == torch.bfloat16 and tuple(y16.shape) == tuple(ref.shape)
    d = (y16.float() - ref.to(torch.bfloat16).float()).abs().max().item()
    assert d <= float(ref.abs().max()) * 2.0 ** -7
    assert (y16 == ref.to(torch.bfloat16)).float().mean().item() > 0.995
    # consumer: 8 -> 4 conv reading the bf16 tensor directly
    w2 = _bf16_exact(rng_arr(rng, (3, 3, 3, 8, 4), 0.2))
    b2 = rng_arr(rng, (4,), 0.1)
    spec2 = ops.ConvSpec(3, 8, 4, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1)
    ref2 = ops.conv_fwd(y16.float(), dev(w2, cuda), dev(b2, cuda), spec2)
    got2 = ops.conv_fwd_small_bf16(y16, dev(w2, cuda), dev(b2, cuda), spec2)
    assert rel_err(got2.cpu().numpy(), ref2.cpu().numpy()) < 2e-5


@pytest.mark.parametrize("ndim,dims,cout,d2s,d2t,roll", [(2, (1, 12, 16), 1600, 5, 1, 0),
                                                          (3, (4, 5, 8), 768, 1, 24, 12),
                                                          (3, (4, 5, 8), 288, 3, 1, 0)])
def test_umma_wide_scatter_head_in_channel_slices(cuda, ndim, dims, cout, d2s, d2t, roll):
    """cout > 256 with a depth_to_space / depth_to_time map: <= 256-channel slices of the
    convolution (cout_total / cout_base) into one destination == the fp32 direct kernel."""
    from sup3r_b200 import ops
    import dataclasses
    rng = np.random.default_rng(cout)
    n = 2
    shape = (n, *dims[1:], 64) if ndim == 2 else (n, *dims, 64)
    k = (3,) * ndim
    x = _bf16_exact(rng_arr(rng, shape))
    w = _bf16_exact(rng_arr(rng, (*k, 64, cout), 0.05))
    b = rng_arr(rng, (cout,), 0.1)
    xd, wd, bd = dev(x, cuda), dev(w, cuda), dev(b, cuda)
    pz = 1 if ndim == 3 else 0
    spec = ops.ConvSpec(ndim, 64, cout, (3 if ndim == 3 else 1, 3, 3), pad_lo=(pz, 1, 1),
                        pad_hi=(pz, 1, 1), pad_mode=1, act=2, alpha=0.2, d2s=d2s, d2t=d2t,
                        t_roll=roll)
    ref = ops.conv_fwd(xd, wd, bd, spec)
    x_hi, _ = ops.pack_act_pad16(xd)
    y = torch.empty_like(ref)
    kd = (1, *dims[1:]) if ndim == 2 else dims
    for cb in range(0, cout, 256):
        nc = min(256, cout - cb)
        w_hi, _ = ops.pack_weights_umma(wd[..., cb:cb + nc].contiguous(), ndim=ndim)
        sp = dataclasses.replace(spec, cout=nc, cout_total=cout, cout_base=cb)
        ops.conv_fwd_umma(x_hi, None, w_hi, None, bd[cb:cb + nc].contiguous(), sp, n, kd, out=y)
    assert rel_err(y.cpu().numpy(), ref.cpu().numpy()) < 2e-5


def test_umma_split_input_channels_pre_activation_residual(cuda):
    """65-channel 2-D conv (64 features + 1 Sup3rConcat exo channel) = tensor-core conv on the 64
    + fp32 conv on the rest, summed BEFORE the LeakyReLU (res_pre_act)."""
    from sup3r_b200 import ops
    import dataclasses
    rng = np.random.default_rng(65)
    n, dims = 3, (1, 12, 20)
    x = _bf16_exact(rng_arr(rng, (n, 12, 20, 65)))
    w = _bf16_exact(rng_arr(rng, (3, 3, 65, 64), 0.05))
    b = rng_arr(rng, (64,), 0.1)
    xd, wd, bd = dev(x, cuda), dev(w, cuda), dev(b, cuda)
    spec = ops.ConvSpec(2, 65, 64, (1, 3, 3), pad_lo=(0, 1, 1), pad_hi=(0, 1, 1), pad_mode=1,
                        act=2, alpha=0.2)
    ref = ops.conv_fwd(xd, wd, bd, spec)
    part = ops.conv_fwd(xd[..., 64:].contiguous(), wd[..., 64:, :].contiguous(), None,
                        dataclasses.replace(spec, cin=1, act=0))
    x_hi, _ = ops.pack_act_pad16(xd[..., :64].contiguous())
    w_hi, _ = ops.pack_weights_umma(wd[..., :64, :].contiguous(), ndim=2)
    y, _, _ = ops.conv_fwd_umma(x_hi, None, w_hi, None, bd,
                                dataclasses.replace(spec, cin=64, res_pre_act=1), n, dims,
                                residual=part)
    assert rel_err(y.cpu().numpy(), ref.cpu().numpy()) < 2e-5


# --------------------------------------------------------------- fp16c (fp16 + e4m3 correction)
def _conv_ref64(x, w, b, alpha=None, res=None, rep=1):
    """float64 torch restatement: reflect-pad-1 3x3x3 conv + bias [+ LeakyReLU] [+ nearest
    repeat along x] [+ residual]"""
    import torch.nn.functional as F
    xc = torch.as_tensor(x, dtype=torch.float64).permute(0, 4, 1, 2, 3)
    xc = F.pad(xc, (1, 1, 1, 1, 1, 1), mode="reflect")
    wc = torch.as_tensor(w, dtype=torch.float64).permute(4, 3, 0, 1, 2)
    y = F.conv3d(xc, wc, torch.as_tensor(b, dtype=torch.float64)).permute(0, 2, 3, 4, 1)
    if alpha is not None:
        y = F.leaky_relu(y, alpha)
    if rep > 1:
        y = torch.repeat_interleave(y, rep, dim=3)
    if res is not None:
        y = y + torch.as_tensor(res, dtype=torch.float64)
    return y.numpy()


def _halo_ok(t, pz=1):
    """REFLECT halo of a padded (n, z+2, y+2, x+2, c) tensor: plane 0 == plane 2, ..."""
    t = t.view(torch.int16) if t.dtype != torch.int16 else t
    ok = bool((t[:, :, 0] == t[:, :, 2]).all() and (t[:, :, -1] == t[:, :, -3]).all()
              and (t[:, :, :, 0] == t[:, :, :, 2]).all() and (t[:, :, :, -1] == t[:, :, :, -3]).all())
    if pz:
        ok = ok and bool((t[:, 0] == t[:, 2]).all() and (t[:, -1] == t[:, -3]).all())
    return ok


def test_fp16c_pack_roundtrip(cuda):
    from sup3r_b200 import ops
    rng = np.random.default_rng(5)
    x = rng_arr(rng, (2, 5, 6, 7, 64), 3.0)
    hi, corr = ops.pack_act_pad16(dev(x, cuda), split=True, fmt=2)
    assert hi.dtype == torch.float16 and corr.shape == hi.shape
    back = ops.unpack_act_pad16(hi, corr, 3, fmt=2).cpu().numpy()
    # hi (11 bits) + e4m3 residue (4 bits): ~2^-15 relative
    assert np.abs(back - x).max() <= np.abs(x).max() * 2.0 ** -14
    assert np.abs(back - x).max() < 0.2 * np.abs(hi[:, 1:-1, 1:-1, 1:-1].float().cpu().numpy() - x).max()
    assert _halo_ok(hi) and _halo_ok(corr)
    # corr row layout: [lo8 x 32 | a8 x 32] per 32-channel half; a8 = e4m3(x)
    row = corr.view(torch.uint8).reshape(*corr.shape[:-1], 2, 2, 32)
    a8 = row[..., 1, :].reshape(*corr.shape[:-1], 64).view(torch.float8_e4m3fn).float()
    a8 = a8[:, 1:-1, 1:-1, 1:-1].cpu().numpy()
    assert np.abs(a8 - x).max() <= np.abs(x).max() * 2.0 ** -4


@pytest.mark.parametrize("halo", [1, 0])
def test_fp16c_pack_is_exact(cuda, halo):
    """The (vectorised) operand packing, bit for bit: hi = fp16(x) on the padded extent (reflect /
    zero halo), correction row = [e4m3((x - hi) 2^11) | e4m3(x)] per 32-channel half."""
    from sup3r_b200 import ops
    rng = np.random.default_rng(8)
    x = rng_arr(rng, (2, 3, 5, 6, 64), 2.0)
    xd = dev(x, cuda)
    hi, corr = ops.pack_act_pad16(xd, split=True, fmt=2, halo=halo)
    xp = xd.permute(0, 4, 1, 2, 3)
    xp = (torch.nn.functional.pad(xp, (1,) * 6, mode="reflect") if halo == 1
          else torch.nn.functional.pad(xp, (1,) * 6))
    xp = xp.permute(0, 2, 3, 4, 1).contiguous()
    want_hi = xp.half()
    assert torch.equal(hi.reshape(want_hi.shape), want_hi)
    row = corr.view(torch.uint8).reshape(*xp.shape[:-1], 2, 2, 32)
    lo8 = row[..., 0, :].reshape(xp.shape).view(torch.float8_e4m3fn).float()
    a8 = row[..., 1, :].reshape(xp.shape).view(torch.float8_e4m3fn).float()
    assert torch.equal(a8, xp.to(torch.float8_e4m3fn).float())
    assert torch.equal(lo8, ((xp - want_hi.float()) * 2048.0).to(torch.float8_e4m3fn).float())
    # a 16-bit pair (bf16 hi + lo) goes through the same kernel
    bh, bl = ops.pack_act_pad16(xd, split=True, fmt=0, halo=halo)
    assert torch.equal(bh.reshape(xp.shape), xp.bfloat16())
    assert torch.equal(bl.reshape(xp.shape), (xp - xp.bfloat16().float()).bfloat16())
    if halo == 0:
        # zero halo two voxels wide == the zero-padded tensor packed with a one-voxel zero halo
        h2, c2 = ops.pack_act_pad16(xd, split=True, fmt=2, halo=0, halo_width=2)
        h1, c1 = ops.pack_act_pad16(ops.pad_fwd(xd, [(0, 0)] + [(1, 1)] * 3 + [(0, 0)], 0),
                                    split=True, fmt=2, halo=0)
        assert h2.shape == h1.shape and torch.equal(h2, h1)
        assert torch.equal(c2.view(torch.uint8), c1.view(torch.uint8))


@pytest.mark.parametrize("nd,cin,cout", [(3, 6, 32), (3, 64, 64), (3, 128, 96), (2, 64, 200),
                                         (3, 32, 6)])
def test_weight_view_packing_equals_packing_of_copies(cuda, nd, cin, cout):
    """s3_pack_weights_umma_view == s3_pack_weights_umma_c of the sliced / zero-padded (forward)
    or flipped + transposed (input gradient) copies it replaces, bit for bit."""
    from sup3r_b200 import ops
    rng = np.random.default_rng(12)
    w = dev(rng_arr(rng, (3,) * nd + (cin, cout), 0.3), cuda)
    wmax = float(w.abs().max())
    pad = torch.nn.functional.pad
    # forward views: 64 input channels from ci0, up to 256 output channels from co0
    for ci0 in range(0, cin, 64):
        for co0 in range(0, cout, 256):
            n_co = min(256, cout - co0)
            ref = w[..., ci0:ci0 + 64, co0:co0 + n_co]
            if ref.shape[-2] < 64:
                ref = pad(ref, (0, 0, 0, 64 - ref.shape[-2]))
            want = ops.pack_weights_umma(ref.contiguous(), ndim=nd, fmt=ops.S3_FMT_FP16C, wmax=wmax)
            got = ops.pack_weights_umma_view(w, ci0, co0, n_co, ndim=nd, wmax=wmax)
            assert torch.equal(got[0], want[0]) and got[2] == want[2]
            assert torch.equal(got[1].view(torch.uint8), want[1].view(torch.uint8))
    # adjoint views: rows = 64 output channels from 64 go, columns = input channels (padded to 16)
    cin_p = (cin + 15) // 16 * 16
    for go in range((cout + 63) // 64):
        wt = w[..., 64 * go:64 * go + 64]
        if wt.shape[-1] < 64:
            wt = pad(wt, (0, 64 - wt.shape[-1]))
        wt = wt.flip(dims=tuple(range(nd))).transpose(-1, -2)
        if cin_p != cin:
            wt = pad(wt, (0, cin_p - cin))
        for c0 in range(0, cin_p, 256):
            c1 = min(cin_p, c0 + 256)
            want = ops.pack_weights_umma(wt[..., c0:c1].contiguous(), ndim=nd,
                                         fmt=ops.S3_FMT_FP16C, wmax=wmax)
            got = ops.pack_weights_umma_view(w, 64 * go, c0, c1 - c0, adjoint=True, ndim=nd,
                                             wmax=wmax)
            assert torch.equal(got[0], want[0]) and got[2] == want[2]
            assert torch.equal(got[1].view(torch.uint8), want[1].view(torch.uint8))


FP16C_CASES = [
    # n, (z, y, x), variant
    (1, (16, 16, 24), "pad16"),          # hot shape: straight-line two-pass MMA role, V4 epilogue
    (2, (8, 16, 40), "pad16"),
    (1, (16, 16, 24), "res16"),          # (hi, corr) residual pair through the TMA epilogue
    (2, (4, 13, 21), "res16"),           # ragged y / x tiles
    (1, (6, 9, 17), "pad16"),            # planes % 4 != 0 -> generic two-pass MMA role
    (1, (16, 16, 24), "f32"),            # fp32 destination (thread-per-row epilogue)
    (1, (8, 16, 24), "res_f32"),
    (1, (8, 16, 8), "rep3"),             # nearest repeat x3 with (hi, corr) output (TMA epilogue,
    (2, (16, 16, 24), "rep2"),           #   one store per replica through 5-D maps)
    (1, (4, 20, 16), "rep2"),            # ragged y tile
    (1, (6, 16, 12), "rep3"),            # x % 8 != 0 / planes % 4 != 0 -> thread-per-row epilogue
]


@pytest.mark.parametrize("n,dims,variant", FP16C_CASES)
def test_umma_conv_fp16c_matches_float64(cuda, n, dims, variant):
    """tcgen05 64->64 3x3x3 conv in the fp16c format (kind::f16 pass + kind::f8f6f4 correction
    pass) against a float64 torch restatement on ARBITRARY fp32 operands: the north-star bound
    is 1e-3 relative; one layer must sit near 2^-15."""
    from sup3r_b200 import ops
    rng = np.random.default_rng(zlib.crc32(repr((n, dims, variant, "c")).encode()))
    x = rng_arr(rng, (n, *dims, 64))
    w = rng_arr(rng, (3, 3, 3, 64, 64), 0.05)
    b = rng_arr(rng, (64,), 0.1)
    rep = int(variant[3:]) if variant.startswith("rep") else 1
    odims = (dims[0], dims[1], dims[2] * rep)
    res = rng_arr(rng, (n, *odims, 64))
    xd, wd, bd, rd = dev(x, cuda), dev(w, cuda), dev(b, cuda), dev(res, cuda)
    has_res = variant.startswith("res")
    spec = ops.ConvSpec(3, 64, 64, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1,
                        act=0 if has_res else 2, alpha=0.2, out_repeat=(1, 1, rep))
    x_hi, x_c = ops.pack_act_pad16(xd, split=True, fmt=2)
    w_hi, w_c, acc_scale = ops.pack_weights_umma(wd, ndim=3, fmt=2)
    r_hi, r_c = ops.pack_act_pad16(rd, split=True, fmt=2)
    res_eff = None
    if has_res:
        res_eff = (res if variant == "res_f32"
                   else ops.unpack_act_pad16(r_hi, r_c, 3, fmt=2).cpu().numpy())
    ref = _conv_ref64(x, w, b, None if has_res else 0.2, res_eff, rep)
    scale = np.abs(ref).max()
    if variant in ("f32", "res_f32"):
        y, _, _ = ops.conv_fwd_umma(x_hi, x_c, w_hi, w_c, bd, spec, n, dims, fmt=2,
                                    acc_scale=acc_scale, residual=rd if has_res else None)
        assert np.abs(y.cpu().numpy() - ref).max() < scale * 1e-4
        return
    _, y_hi, y_c = ops.conv_fwd_umma(x_hi, x_c, w_hi, w_c, bd, spec, n, dims, want_f32=False,
                                     want_pad16=True, want_lo=True, fmt=2, acc_scale=acc_scale,
                                     res_hi=r_hi if has_res else None,
                                     res_lo=r_c if has_res else None)
    got = ops.unpack_act_pad16(y_hi, y_c, 3, fmt=2).cpu().numpy()
    err = np.abs(got - ref).max() / scale
    assert err < 1.5e-4, err          # 2^-14 storage + 2^-15 operands
    # hi alone carries fp16 rounding (2^-12): the correction rows must make a visible difference
    err_hi = np.abs(y_hi[:, 1:-1, 1:-1, 1:-1].float().cpu().numpy() - ref).max() / scale
    assert err < 0.5 * err_hi, (err, err_hi)
    assert _halo_ok(y_hi) and _halo_ok(y_c)
    # the a8 copy in the corr rows is e4m3(value)
    row = y_c.view(torch.uint8).reshape(*y_c.shape[:-1], 2, 2, 32)
    a8 = row[..., 1, :].reshape(*y_c.shape[:-1], 64).view(torch.float8_e4m3fn).float()
    a8 = a8[:, 1:-1, 1:-1, 1:-1].cpu().numpy()
    assert np.abs(a8 - ref).max() <= scale * 2.0 ** -4 * 1.01


def test_umma_conv_fp16c_2d_and_head(cuda):
    """fp16c on the tile kernel: a 2-D 64->64 conv with (hi, corr) output and a 3-D 64->200 head
    with fused 5x depth_to_space into an f32 tensor."""
    from sup3r_b200 import ops
    import torch.nn.functional as F
    rng = np.random.default_rng(11)
    # 2-D
    n, dims = 3, (1, 12, 20)
    x = rng_arr(rng, (n, 12, 20, 64))
    w = rng_arr(rng, (3, 3, 64, 64), 0.05)
    b = rng_arr(rng, (64,), 0.1)
    xc = F.pad(torch.as_tensor(x, dtype=torch.float64).permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect")
    ref = F.conv2d(xc, torch.as_tensor(w, dtype=torch.float64).permute(3, 2, 0, 1),
                   torch.as_tensor(b, dtype=torch.float64)).permute(0, 2, 3, 1)
    ref = F.leaky_relu(ref, 0.2).numpy()
    spec = ops.ConvSpec(2, 64, 64, (1, 3, 3), pad_lo=(0, 1, 1), pad_hi=(0, 1, 1), pad_mode=1,
                        act=2, alpha=0.2)
    x_hi, x_c = ops.pack_act_pad16(dev(x, cuda), split=True, fmt=2)
    w_hi, w_c, sc = ops.pack_weights_umma(dev(w, cuda), ndim=2, fmt=2)
    _, y_hi, y_c = ops.conv_fwd_umma(x_hi, x_c, w_hi, w_c, dev(b, cuda), spec, n, dims,
                                     want_f32=False, want_pad16=True, want_lo=True, fmt=2,
                                     acc_scale=sc)
    got = ops.unpack_act_pad16(y_hi, y_c, 2, fmt=2).cpu().numpy()
    assert np.abs(got - ref).max() < np.abs(ref).max() * 1.5e-4
    assert _halo_ok(y_hi, pz=0) and _halo_ok(y_c, pz=0)
    # 3-D head 64 -> 200, depth_to_space 5
    n, dims = 1, (6, 8, 16)
    x = rng_arr(rng, (n, *dims, 64))
    w = rng_arr(rng, (3, 3, 3, 64, 200), 0.05)
    b = rng_arr(rng, (200,), 0.1)
    conv = _conv_ref64(x, w, b, 0.2)
    ref = np.stack([L.depth_to_space(conv[:, :, :, i], 5) for i in range(dims[2])], axis=3)
    spec = ops.ConvSpec(3, 64, 200, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1), pad_mode=1,
                        act=2, alpha=0.2, d2s=5)
    x_hi, x_c = ops.pack_act_pad16(dev(x, cuda), split=True, fmt=2)
    w_hi, w_c, sc = ops.pack_weights_umma(dev(w, cuda), ndim=3, fmt=2)
    y, _, _ = ops.conv_fwd_umma(x_hi, x_c, w_hi, w_c, dev(b, cuda), spec, n, dims, fmt=2,
                                acc_scale=sc)
    assert y.shape == ref.shape
    assert np.abs(y.cpu().numpy() - ref).max() < np.abs(ref).max() * 1e-4


# ------------------------------------------------------------------ tcgen05 weight gradient
@pytest.mark.parametrize("n,dims,cin,halo", [(1, (4, 16, 8), 64, 1), (2, (5, 20, 13), 64, 2),
                                              (1, (2, 7, 30), 6, 1), (3, (9, 33, 17), 64, 2)])
def test_umma_wgrad_matches_float64(cuda, n, dims, cin, halo):
    """dW = sum_v x_pad[v + tap] (x) dy[v] with voxels as the K dimension of MN-major tcgen05
    operands (conv_wgrad_umma.cu) against float64 autograd of the same fp16-rounded operands
    (tight) and of the unrounded ones (the 2e-3 gradient bound)."""
    import torch.nn.functional as F
    from sup3r_b200 import ops
    rng = np.random.default_rng(n * 100 + cin)
    x = rng.standard_normal((n, *dims, cin)).astype(np.float32)
    dy = (rng.standard_normal((n, *dims, 64)) * 0.5).astype(np.float32)
    xd = torch.as_tensor(x, device=cuda)
    x_hi, _ = ops.pack_act_pad16(torch.nn.functional.pad(xd, (0, 64 - cin)), split=False,
                                 fmt=ops.S3_FMT_FP16)
    gd = torch.as_tensor(dy, device=cuda)
    if halo == 2:
        gd = ops.pad_fwd(gd, [(0, 0), (1, 1), (1, 1), (1, 1), (0, 0)], 0)
    g_hi, _ = ops.pack_act_pad16(gd, split=False, fmt=ops.S3_FMT_FP16, halo=0)
    got = ops.conv_wgrad_umma(x_hi, g_hi, halo, n, dims, cin).cpu().numpy().astype(np.float64)
    assert got.shape == (3, 3, 3, cin, 64)

    def ref(xa, ga):
        xc = torch.tensor(xa, dtype=torch.float64).permute(0, 4, 1, 2, 3)
        xp = F.pad(xc, (1, 1, 1, 1, 1, 1), mode="reflect")
        w = torch.zeros((64, cin, 3, 3, 3), dtype=torch.float64, requires_grad=True)
        y = F.conv3d(xp, w)
        (dw,) = torch.autograd.grad(y, w, torch.tensor(ga, dtype=torch.float64).permute(0, 4, 1, 2, 3))
        return dw.permute(2, 3, 4, 1, 0).numpy()

    r16 = ref(x.astype(np.float16).astype(np.float64), dy.astype(np.float16).astype(np.float64))
    assert np.abs(got - r16).max() < 2e-5 * np.abs(r16).max()
    rex = ref(x, dy)
    assert np.abs(got - rex).max() < 2e-3 * np.abs(rex).max()
