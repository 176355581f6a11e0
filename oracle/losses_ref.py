"""CPU restatement (numpy, float64) of the content losses in sup3r/utilities/loss_metrics.py.

TEST INFRASTRUCTURE ONLY: imported by tests/ as the checker of sup3r_b200.loss_metrics, never by
the product.  Each function cites the reference lines it follows; keras MeanSquaredError /
MeanAbsoluteError reduce to the global mean for the equally-shaped tensors used here.
PINNED: ``tests/golden/losses.json`` holds the values of the REAL reference classes
(``tools/make_golden_losses.py`` execs sup3r/utilities/loss_metrics.py with a numpy-backed ``tf``
stub); ``tests/test_losses_golden.py`` holds this file to them at 1e-12.  Also checked against
the reference's own known-answer identities (tests/utilities/test_loss_metrics.py:174-309):
np.gradient equality of the material derivative, LowResLoss == MSE without coarsening,
LowResLoss on pre-coarsened fields, extremes dominated by single spikes.
"""
import numpy as np


def mse(a, b):
    return float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))


def mae(a, b):
    return float(np.mean(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


def derivative(x, axis):
    """loss_metrics.py:12-59 (== np.gradient, unit spacing)."""
    if axis not in (1, 2, 3):
        raise ValueError(f"_derivative received axis={axis}")
    return np.gradient(np.asarray(x, np.float64), axis=axis)


def exp_loss(x1, x2):
    """loss_metrics.py:98-118"""
    return float(np.mean(1 - np.exp(-((np.asarray(x1, np.float64) - x2) ** 2))))


def gaussian_kernel(x1, x2, sigma=1.0):
    """loss_metrics.py:62-95"""
    d = np.expand_dims(np.asarray(x1, np.float64), 1) - np.asarray(x2, np.float64)
    return np.exp(-0.5 * np.sum(d ** 2, axis=-1) / sigma ** 2)


def mmd_loss(x1, x2, sigma=1.0):
    """loss_metrics.py:121-147"""
    return float(np.mean(gaussian_kernel(x1, x1, sigma)) + np.mean(gaussian_kernel(x2, x2, sigma))
                 - np.mean(2 * gaussian_kernel(x1, x2, sigma)))


def compute_md(x, fidx):
    """loss_metrics.py:162-188"""
    x = np.asarray(x, np.float64)
    u, v = 2 * (fidx // 2), 2 * (fidx // 2) + 1
    out = derivative(x[..., fidx], 3)
    out = out + x[..., u] * derivative(x[..., fidx], 1)
    out = out + x[..., v] * derivative(x[..., fidx], 2)
    return out


def material_derivative_loss(x1, x2):
    """loss_metrics.py:190-225"""
    assert np.ndim(x1) == 5 and np.ndim(x2) == 5
    hh = np.shape(x1)[-1] // 2
    a = np.stack([compute_md(x1, i) for i in range(0, 2 * hh, 2)])
    b = np.stack([compute_md(x2, i) for i in range(0, 2 * hh, 2)])
    return mae(a, b)


def spatial_derivative_loss(x1, x2):
    """loss_metrics.py:233-260"""
    return mae(derivative(x1, 1) + derivative(x1, 2), derivative(x2, 1) + derivative(x2, 2))


def temporal_derivative_loss(x1, x2):
    """loss_metrics.py:268-294"""
    return mae(derivative(x1, 3), derivative(x2, 3))


def coarse_mse_loss(x1, x2):
    """loss_metrics.py:302-322"""
    return mse(np.mean(x1, axis=(1, 2)), np.mean(x2, axis=(1, 2)))


def spatial_extremes_loss(x1, x2):
    """loss_metrics.py:331-357"""
    return (mae(np.min(x1, axis=(1, 2)), np.min(x2, axis=(1, 2)))
            + mae(np.max(x1, axis=(1, 2)), np.max(x2, axis=(1, 2)))) / 2


def temporal_extremes_loss(x1, x2):
    """loss_metrics.py:366-392"""
    return (mae(np.min(x1, axis=3), np.min(x2, axis=3))
            + mae(np.max(x1, axis=3), np.max(x2, axis=3))) / 2


def _fft_feature(x, axes):
    x = np.asarray(x, np.float64)
    w = 1.0
    for k, ax in enumerate(axes):
        shp = [1] * x.ndim
        shp[ax] = x.shape[ax]
        w = w * (np.arange(x.shape[ax], dtype=np.float64) ** 2).reshape(shp)
    return np.log(1 + w * np.abs(np.fft.fftn(x, axes=axes)))


def spatial_fft_loss(x1, x2):
    """loss_metrics.py:395-437"""
    return mae(_fft_feature(x1, (1, 2)), _fft_feature(x2, (1, 2)))


def spatiotemporal_fft_loss(x1, x2):
    """loss_metrics.py:440-485"""
    return mae(_fft_feature(x1, (1, 2, 3)), _fft_feature(x2, (1, 2, 3)))


def low_res_loss(x1, x2, s_enhance=1, t_enhance=1, t_method="average",
                 tf_loss="MeanSquaredError", ex_loss=None):
    """loss_metrics.py:488-638"""
    x1, x2 = np.asarray(x1, np.float64), np.asarray(x2, np.float64)
    ex = 0.0
    if ex_loss is not None:
        ex = {"SpatialExtremesLoss": spatial_extremes_loss,
              "TemporalExtremesLoss": temporal_extremes_loss}[ex_loss](x1, x2)

    def s_coarsen(t):
        s = s_enhance
        shp = t.shape
        t = t.reshape((shp[0], shp[1] // s, s, shp[2] // s, s) + shp[3:])
        return t.sum(axis=(2, 4)) / s ** 2

    if s_enhance > 1:
        x1, x2 = s_coarsen(x1), s_coarsen(x2)
    if t_enhance > 1 and t_method.casefold() == "average":
        def t_avg(t):
            n, a, b, _, f = t.shape
            return t.reshape(n, a, b, -1, t_enhance, f).sum(axis=4) / t_enhance
        x1, x2 = t_avg(x1), t_avg(x2)
    if t_enhance > 1 and t_method.casefold() == "subsample":
        x1, x2 = x1[:, :, :, ::t_enhance, :], x2[:, :, :, ::t_enhance, :]
    return {"MeanSquaredError": mse, "MeanAbsoluteError": mae}[tf_loss](x1, x2) + ex


def sliced_wasserstein_loss(x1, x2, proj):
    """loss_metrics.py:754-793 with the (n_projections, H*W*T) projection matrix given (the
    reference draws it with tf.random.normal and l2-normalises the rows)."""
    x1, x2 = np.asarray(x1, np.float64), np.asarray(x2, np.float64)
    b, c = x1.shape[0], x1.shape[-1]
    p1 = np.asarray(proj, np.float64) @ x1.reshape(b, -1, c)      # (B, P, C)
    p2 = np.asarray(proj, np.float64) @ x2.reshape(b, -1, c)
    return float(np.mean((np.sort(p1, axis=1) - np.sort(p2, axis=1)) ** 2))

