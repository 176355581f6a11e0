"""Pins for the CPU oracle (numpy restatement, oracle/layers_ref.py).

The convolution arithmetic of the reference lives in TensorFlow / phygnn, which are not
importable here, and the reference tests hold no golden outputs for it: parity is UNPINNED for
conv arithmetic.  What can be pinned is pinned here: (1) the reference's own shape tables
(tests/training/test_load_configs.py), (2) an independent float64 implementation of every layer
(torch CPU ops, oracle/torch_ref.py) on seeded inputs, (3) the algebraic identities the fused
kernels rely on (pad-3/conv/crop-2 == reflect-1 conv; Conv2DTranspose == flipped conv),
(4) TensorFlow's documented conventions for SAME padding and depth_to_space on hand-computed
examples."""
import numpy as np
import pytest
import torch

from oracle import layers_ref as L
from oracle.torch_ref import TorchRefNet
from sup3r_b200 import configs as C


def build(hl, in_shape, seed=0):
    layers = L.build_layers(hl)
    L.build_weights(layers, in_shape, seed=seed)
    return layers


CASES = [
    ("st_5x_12x", C.spatiotemporal_generator(4, 5, (2, 2, 3), head_filters=200), (1, 5, 6, 4, 4), None),
    ("st_3x_4x", C.spatiotemporal_generator(2, 3, (2, 2)), (2, 5, 5, 4, 2), None),
    ("s_2x", C.spatial_generator(2, (2,)), (3, 7, 8, 2), None),
    ("s_10x", C.spatial_generator(2, (2, 5), n_blocks=2), (1, 6, 6, 2), None),
    ("cc_trh_d2t", C.sup3rcc_temporal_d2t_generator(2, 24, 12, n_blocks=2), (1, 5, 5, 4, 2), None),
    ("cc_wind_exo", C.sup3rcc_spatial_generator(6, 5, 4, exo="topography"), (2, 6, 6, 6),
     ("topography", (2, 30, 30, 1))),
    ("disc_st_same", C.discriminator(3, "same", (64, 32)), (2, 9, 9, 10, 3), None),
    ("disc_s_valid", C.discriminator(2, "valid", (32,)), (1, 64, 64, 2), None),
]


@pytest.mark.parametrize("name,hl,shape,exo", CASES, ids=[c[0] for c in CASES])
def test_numpy_oracle_agrees_with_torch_restatement(name, hl, shape, exo):
    rng = np.random.default_rng(3)
    layers = build(hl, shape)
    x = rng.standard_normal(shape)
    exo_d = {exo[0]: rng.standard_normal(exo[1])} if exo else None
    y = L.run_layers(layers, x.astype(np.float64), exo_d)
    net = TorchRefNet(hl, L.get_weights(layers), dtype=torch.float64)
    yt = net(torch.tensor(x), exo_d).numpy()
    assert y.shape == yt.shape
    assert np.abs(y - yt).max() <= 1e-9 * max(np.abs(yt).max(), 1.0)


def test_shape_tables_of_reference_tests():
    """tests/training/test_load_configs.py: (n, 7, 7, [4], f) ones -> enhanced shapes."""
    for s, t, f, tm in [(3, 4, 2, (2, 2)), (2, 12, 14, (2, 2, 3)), (4, 24, 3, (2, 2, 2, 3))]:
        hl = C.spatiotemporal_generator(f, s, tm)
        layers = L.build_layers(hl)
        shp = (1, 7, 7, 4, f)
        for lyr in layers:
            shp = L.out_shape(lyr, shp)
        assert shp == (1, 7 * s, 7 * s, 4 * t, f)
    layers = build(C.spatial_generator(2, (2,)), (4, 10, 10, 2))
    y = L.run_layers(layers, np.ones((4, 10, 10, 2), np.float32))
    assert y.shape == (4, 20, 20, 2) and y.dtype == np.float32


def test_fusion_identities():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 6, 7, 5, 3))
    w = rng.standard_normal((3, 3, 3, 3, 4))
    b = rng.standard_normal(4)
    lit = L.crop_nd(L.conv_nd(L.tf_pad(x, [[0, 0], [3, 3], [3, 3], [3, 3], [0, 0]], "REFLECT"),
                              w, b), 2)
    fused = L.conv_nd(L.tf_pad(x, [[0, 0], [1, 1], [1, 1], [1, 1], [0, 0]], "REFLECT"), w, b)
    assert np.abs(lit - fused).max() < 1e-12
    x2 = rng.standard_normal((2, 7, 8, 5))
    wt = rng.standard_normal((3, 3, 6, 5))
    lit = L.crop_nd(L.conv_transpose_nd(L.tf_pad(x2, [[0, 0], [3, 3], [3, 3], [0, 0]], "REFLECT"),
                                        wt), 4)
    wf = wt[::-1, ::-1].transpose(0, 1, 3, 2)
    fused = L.conv_nd(L.tf_pad(x2, [[0, 0], [1, 1], [1, 1], [0, 0]], "REFLECT"), wf)
    assert np.abs(lit - fused).max() < 1e-12


def test_tf_conventions_hand_computed():
    # SAME padding, stride 2: in 5, k 3 -> out 3, total pad 2 -> (1, 1); in 6 -> out 3, total 1
    # -> (0, 1) (the odd element goes AFTER)
    assert L.same_pads(5, 3, 2) == (1, 1) and L.same_pads(6, 3, 2) == (0, 1)
    assert L.same_pads(7, 3, 1) == (1, 1)
    # depth_to_space DCR: out[h*r+i, w*r+j, c] = in[h, w, (i*r+j)*C' + c]
    x = np.arange(8, dtype=np.float32).reshape(1, 1, 1, 8)
    y = L.depth_to_space(x, 2)
    assert y.shape == (1, 2, 2, 2)
    assert y[0, 0, 0].tolist() == [0, 1] and y[0, 0, 1].tolist() == [2, 3]
    assert y[0, 1, 0].tolist() == [4, 5] and y[0, 1, 1].tolist() == [6, 7]
    # depth_to_time = row-major reshape then roll
    x = np.arange(12, dtype=np.float32).reshape(1, 1, 1, 2, 6)
    y = L.spatiotemporal_expansion(x, 1, 3, "depth_to_time", 1)
    assert y.shape == (1, 1, 1, 6, 2)
    assert y[0, 0, 0, :, 0].tolist() == [10, 0, 2, 4, 6, 8]
    # REFLECT excludes the edge, SYMMETRIC repeats it
    a = np.arange(4.0).reshape(1, 4, 1)
    assert L.tf_pad(a, [[0, 0], [2, 1], [0, 0]], "REFLECT")[0, :, 0].tolist() == [2, 1, 0, 1, 2, 3, 2]
    assert L.tf_pad(a, [[0, 0], [2, 1], [0, 0]], "SYMMETRIC")[0, :, 0].tolist() == [1, 0, 0, 1, 2, 3, 3]
    with pytest.raises(ValueError):
        L.tf_pad(a, [[0, 0], [4, 0], [0, 0]], "REFLECT")


def test_skip_connection_cache_semantics():
    s = L.SkipConnection("a")
    x = np.ones((1, 2, 2, 1))
    assert s(x) is x
    assert np.all(s(2 * x) == 3)
    assert s(x) is x  # cache cleared after the add


def test_loss_identities():
    """LowResLoss-style identity of tests/utilities/test_loss_metrics.py:174-260 on the numpy
    side: coarsening then MSE equals MSE of coarsened tensors; relativistic loss symmetric."""
    from oracle.torch_ref import disc_loss
    from sup3r_b200.utilities import spatial_coarsening, temporal_coarsening
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((2, 8, 8, 6, 2)), rng.standard_normal((2, 8, 8, 6, 2))
    ca = temporal_coarsening(spatial_coarsening(a, 2), 3, "average")
    cb = temporal_coarsening(spatial_coarsening(b, 2), 3, "average")
    assert ca.shape == (2, 4, 4, 2, 2)
    assert np.isclose(((ca - cb) ** 2).mean(),
                      ((temporal_coarsening(spatial_coarsening(a - b, 2), 3, "average")) ** 2).mean())
    t, g = torch.randn(5, 1), torch.randn(5, 1)
    assert disc_loss(t, g) > 0
    c = torch.full((5, 1), 0.7)
    assert torch.isclose(disc_loss(c, c), torch.log(torch.tensor(2.0)))


@pytest.mark.parametrize("nd,strides,padding", [(3, 1, "valid"), (2, 1, "valid"), (3, 2, "valid"),
                                                 (2, 2, "same"), (3, 1, "same")])
def test_conv_oracle_agrees_with_scipy_correlate(nd, strides, padding):
    """Third independent statement of the Keras ConvND arithmetic (cross-correlation, kernel
    (k..., cin, cout), 'same' = TF end-biased zero padding): scipy.signal.correlate per channel
    pair, strided by subsampling the stride-1 result."""
    from scipy.signal import correlate
    rng = np.random.default_rng(nd * 10 + strides)
    sp = (6, 7, 5)[:nd]
    cin, cout = 3, 2
    x = rng.standard_normal((2, *sp, cin))
    w = rng.standard_normal((3,) * nd + (cin, cout))
    b = rng.standard_normal(cout)
    got = L.conv_nd(x, w, b, strides, padding)
    xp = x
    if padding == "same":
        pads = [(0, 0)] + [L.same_pads(sp[d], 3, strides) for d in range(nd)] + [(0, 0)]
        xp = np.pad(x, pads)
    ref = np.zeros_like(got)
    sl = (slice(None, None, strides),) * nd
    for n in range(x.shape[0]):
        for co in range(cout):
            acc = 0.0
            for ci in range(cin):
                acc = acc + correlate(xp[n, ..., ci], w[..., ci, co], mode="valid")[sl]
            ref[n, ..., co] = acc + b[co]
    assert np.allclose(got, ref, atol=1e-12)


def test_published_tensorflow_api_examples():
    """Known answers printed in the TensorFlow API documentation of the ops the generator is
    built from (third-party published vectors; TensorFlow itself is not installed)."""
    # tf.nn.depth_to_space (block_size 2, NHWC): the three examples of its docstring
    x = np.array([[[[1, 2, 3, 4]]]], np.float32)
    assert L.depth_to_space(x, 2).tolist() == [[[[1], [2]], [[3], [4]]]]
    x = np.array([[[[1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12]]]], np.float32)
    assert L.depth_to_space(x, 2).tolist() == [[[[1, 2, 3], [4, 5, 6]], [[7, 8, 9], [10, 11, 12]]]]
    x = np.array([[[[1, 2, 3, 4], [5, 6, 7, 8]], [[9, 10, 11, 12], [13, 14, 15, 16]]]], np.float32)
    assert L.depth_to_space(x, 2).tolist() == [[[[1], [2], [5], [6]], [[3], [4], [7], [8]],
                                                [[9], [10], [13], [14]], [[11], [12], [15], [16]]]]
    # tf.pad: t = [[1, 2, 3], [4, 5, 6]], paddings = [[1, 1], [2, 2]]
    t = np.array([[1, 2, 3], [4, 5, 6]], np.float32)
    assert L.tf_pad(t, [[1, 1], [2, 2]], "REFLECT").tolist() == [
        [6, 5, 4, 5, 6, 5, 4], [3, 2, 1, 2, 3, 2, 1], [6, 5, 4, 5, 6, 5, 4], [3, 2, 1, 2, 3, 2, 1]]
    assert L.tf_pad(t, [[1, 1], [2, 2]], "SYMMETRIC").tolist() == [
        [2, 1, 1, 2, 3, 3, 2], [2, 1, 1, 2, 3, 3, 2], [5, 4, 4, 5, 6, 6, 5], [5, 4, 4, 5, 6, 6, 5]]
    assert L.tf_pad(t, [[1, 1], [2, 2]], "CONSTANT").tolist() == [
        [0, 0, 0, 0, 0, 0, 0], [0, 0, 1, 2, 3, 0, 0], [0, 0, 4, 5, 6, 0, 0], [0, 0, 0, 0, 0, 0, 0]]


def test_published_tensorflow_conv_and_loss_examples():
    """More known answers from the TensorFlow / Keras API documentation: the ``tf.nn.conv2d``
    example (cross-correlation, kernel laid out (kh, kw, cin, cout), VALID), ``LeakyReLU``,
    ``sigmoid_cross_entropy_with_logits``, keras ``MeanSquaredError`` / ``MeanAbsoluteError``,
    ``tf.roll`` (the shift of the depth-to-time expansion) and the first keras ``Adam`` step."""
    import torch
    from oracle import losses_ref, torch_ref
    x_in = np.array([[[[2], [1], [2], [0], [1]], [[1], [3], [2], [2], [3]], [[1], [1], [3], [3], [0]],
                      [[2], [2], [0], [1], [1]], [[0], [0], [3], [1], [2]]]], np.float64)
    kernel_in = np.array([[[[2, 0.1]], [[3, 0.2]]], [[[0, 0.3]], [[1, 0.4]]]])
    want = np.array([[[[10., 1.9], [10., 2.2], [6., 1.6], [6., 2.]],
                      [[12., 1.4], [15., 2.2], [13., 2.7], [13., 1.7]],
                      [[7., 1.7], [11., 1.3], [16., 1.3], [7., 1.]],
                      [[10., 0.6], [7., 1.4], [4., 1.5], [7., 1.4]]]])
    np.testing.assert_allclose(L.conv_nd(x_in, kernel_in, None, 1, "valid"), want, atol=1e-12)
    # the same through the Conv2D layer objects of both oracles (keras weight order)
    hl = [{"class": "Conv2D", "filters": 2, "kernel_size": 2, "strides": 1}]
    layers = L.build_layers(hl)
    L.build_weights(layers, x_in.shape, seed=0)
    L.set_weights(layers, [kernel_in, np.zeros(2)])
    np.testing.assert_allclose(L.run_layers(layers, x_in), want, atol=1e-12)
    net = torch_ref.TorchRefNet(hl, [kernel_in, np.zeros(2)], torch.float64)
    np.testing.assert_allclose(net(torch.from_numpy(x_in)).numpy(), want, atol=1e-12)
    # tf.keras.layers.LeakyReLU: default alpha 0.3; alpha=0.1
    v = np.array([-3.0, -1.0, 0.0, 2.0])
    np.testing.assert_allclose(L.LeakyReLU()(v), [-0.9, -0.3, 0.0, 2.0], atol=1e-12)
    np.testing.assert_allclose(L.LeakyReLU(alpha=0.1)(v), [-0.3, -0.1, 0.0, 2.0], atol=1e-12)
    # tf.nn.sigmoid_cross_entropy_with_logits
    logits = torch.tensor([1., -1., 0., 1., -1., 0., 0.], dtype=torch.float64)
    labels = torch.tensor([0., 0., 0., 1., 1., 1., 0.5], dtype=torch.float64)
    bce = torch.nn.functional.binary_cross_entropy_with_logits(logits, labels, reduction="none")
    np.testing.assert_allclose(bce.numpy(), [1.3132616, 0.3132617, 0.6931472, 0.3132617, 1.3132616,
                                             0.6931472, 0.6931472], atol=1e-7)
    # (the relativistic loss of the oracle is that function on [true - mean(gen), gen - mean(true)])
    t, g = torch.tensor([[1.0], [2.0]]), torch.tensor([[0.5], [-0.5]])
    lg = torch.cat([t - g.mean(), g - t.mean()])
    lb = torch.cat([torch.ones_like(t), torch.zeros_like(g)])
    assert float(torch_ref.disc_loss(t, g)) == pytest.approx(float(
        torch.nn.functional.binary_cross_entropy_with_logits(lg, lb)), rel=1e-12)
    # tf.keras.losses.MeanSquaredError / MeanAbsoluteError
    y_true, y_pred = np.array([[0., 1.], [0., 0.]]), np.array([[1., 1.], [1., 0.]])
    assert losses_ref.mse(y_pred, y_true) == 0.5 and losses_ref.mae(y_pred, y_true) == 0.5
    # tf.roll
    assert np.roll(np.arange(5.0), 2).tolist() == [3., 4., 0., 1., 2.]
    x = np.arange(12, dtype=np.float64).reshape(1, 1, 1, 2, 6)
    rolled = L.spatiotemporal_expansion(x, 1, 3, "depth_to_time", 2)[0, 0, 0, :, 0]
    assert rolled.tolist() == np.roll(np.arange(0.0, 12.0, 2.0), 2).tolist()
    # tf.repeat(..., axis) repeats element-wise ([1, 1, 2, 2], not a tile): the "nearest" expansion
    x = np.array([1.0, 2.0, 3.0]).reshape(1, 1, 1, 3, 1)
    assert L.spatiotemporal_expansion(x, 1, 2, "nearest", 0)[0, 0, 0, :, 0].tolist() == \
        [1., 1., 2., 2., 3., 3.]
    # tf.keras.optimizers.Adam(learning_rate=0.1) on loss = var ** 2 / 2 from var = 10: one step
    # gives 9.9 (the first Adam step has size lr whatever the gradient)
    p, grad, lr, b1, b2, eps = 10.0, 10.0, 0.1, 0.9, 0.999, 1e-7
    m, vv = (1 - b1) * grad, (1 - b2) * grad * grad
    p -= lr * np.sqrt(1 - b2) / (1 - b1) * m / (np.sqrt(vv) + eps)
    assert p == pytest.approx(9.9, abs=1e-6)
