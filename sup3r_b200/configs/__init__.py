"""Programmatic builders for the generator / discriminator ``hidden_layers`` configs.

The reference ships its networks as JSON files (sup3r/configs/{spatial,spatiotemporal,sup3rcc});
users load those files unchanged (``Sup3rGan('gen.json', 'disc.json')``).  These builders
produce the same layer lists from a handful of parameters so that tests and benchmarks do not
depend on the reference tree being present (tests/test_configs.py checks them for equality
with the reference files when ``/root/reference`` exists)."""
from __future__ import annotations


def _pad(nd, p=3):
    return {"class": "FlexiblePadding",
            "paddings": [[0, 0]] + [[p, p]] * nd + [[0, 0]], "mode": "REFLECT"}


def _conv_block(nd, filters, act=True, expansion=None, crop=2, pad=3):
    """pad -> ConvND -> crop [-> LeakyReLU] [-> expansion] (sup3rcc / spatiotemporal style)."""
    out = [_pad(nd, pad),
           {"class": f"Conv{nd}D", "filters": filters, "kernel_size": 3, "strides": 1},
           {"class": f"Cropping{nd}D", "cropping": crop}]
    if act:
        out.append({"alpha": 0.2, "class": "LeakyReLU"})
    if expansion is not None:
        out.append(expansion)
    return out


def _res_blocks(nd, n, name, filters=64):
    return {"n": n, "repeat": [{"class": "SkipConnection", "name": name},
                               *_conv_block(nd, filters, act=True),
                               *_conv_block(nd, filters, act=False),
                               {"class": "SkipConnection", "name": name}]}


def spatiotemporal_generator(n_out, s_enhance, t_mults=(2, 2), head_filters=None, n_blocks=16,
                             filters=64):
    """sup3r/configs/spatiotemporal/gen_*: nearest temporal expansions in the stem, 16 residual
    blocks with a big skip, head conv -> per-time-slice depth_to_space -> LeakyReLU -> output.
    ``t_mults``: temporal multipliers of the stem blocks (runs of 2s are emitted as a repeat)."""
    if head_filters is None:
        head_filters = 72 if s_enhance in (2, 3) else 8 * s_enhance * s_enhance
    hl = []
    t_mults = list(t_mults)
    n2 = 0
    while n2 < len(t_mults) and t_mults[n2] == 2:
        n2 += 1
    if n2:
        hl.append({"n": n2, "repeat": _conv_block(3, filters, expansion={
            "class": "SpatioTemporalExpansion", "temporal_mult": 2,
            "temporal_method": "nearest"})})
    for m in t_mults[n2:]:
        hl += _conv_block(3, filters, expansion={"class": "SpatioTemporalExpansion",
                                                 "temporal_mult": m,
                                                 "temporal_method": "nearest"})
    hl.append({"class": "SkipConnection", "name": "a"})
    hl.append(_res_blocks(3, n_blocks, "b", filters))
    hl += _conv_block(3, filters, act=False)
    hl.append({"class": "SkipConnection", "name": "a"})
    hl += _conv_block(3, head_filters, act=False)
    hl.append({"class": "SpatioTemporalExpansion", "spatial_mult": s_enhance})
    hl.append({"alpha": 0.2, "class": "LeakyReLU"})
    hl += _conv_block(3, n_out, act=False)
    return hl


def _convt_block(filters, act):
    return [_pad(2, 3),
            {"class": "Conv2DTranspose", "filters": filters, "kernel_size": 3, "strides": 1,
             "activation": "relu" if act else None},
            {"class": "Cropping2D", "cropping": 4}]


def spatial_generator(n_out, s_mults=(2,), n_blocks=16, filters=64):
    """sup3r/configs/spatial/gen_*: Conv2DTranspose blocks with fused relu, one
    ``SpatialExpansion`` stage per entry of ``s_mults`` (head filters = 64 * mult^2)."""
    hl = [*_convt_block(filters, True), {"class": "SkipConnection", "name": "a"},
          {"n": n_blocks, "repeat": [{"class": "SkipConnection", "name": "b"},
                                     *_convt_block(filters, True), *_convt_block(filters, False),
                                     {"class": "SkipConnection", "name": "b"}]},
          *_convt_block(filters, False), {"class": "SkipConnection", "name": "a"}]
    for m in s_mults:
        hl += _convt_block(filters * m * m, False)
        hl.append({"class": "SpatialExpansion", "spatial_mult": m})
        hl.append({"class": "Activation", "activation": "relu"})
    hl += _convt_block(n_out, False)
    return hl


def discriminator(nd, padding="valid", dense=(1024,)):
    """sup3r/configs/{spatial,spatiotemporal}/disc.json (padding 'valid') and the 'same'-padded
    variants the reference tests use (tests/data/config_disc_{s,st}_test.json)."""
    hl = []
    for f in (32, 64, 128, 256):
        for s in (1, 2):
            hl += [{"class": f"Conv{nd}D", "filters": f, "kernel_size": 3, "padding": padding,
                    "strides": s}, {"alpha": 0.2, "class": "LeakyReLU"}]
    hl.append({"class": "Flatten"})
    for u in dense:
        hl += [{"class": "Dense", "units": u}, {"alpha": 0.2, "class": "LeakyReLU"}]
    hl.append({"class": "Dense", "units": 1})
    return hl


def sup3rcc_spatial_generator(n_out, s_enhance=5, n_blocks=16, exo=None, filters=64):
    """sup3r/configs/sup3rcc/gen_{solar,wind}_5x_1x_*: Conv2D residual stack, 5x depth_to_space;
    with ``exo`` (e.g. 'topography') a Sup3rConcat and a second half-depth residual stack."""
    half = n_blocks // 2 if exo else n_blocks
    sfx = "_1" if exo else ""
    hl = [*_conv_block(2, filters), {"class": "SkipConnection", "name": f"big_skip{sfx}"},
          _res_blocks(2, half, f"small_skip{sfx}", filters), *_conv_block(2, filters, act=False),
          {"class": "SkipConnection", "name": f"big_skip{sfx}"},
          *_conv_block(2, filters * s_enhance * s_enhance, act=False),
          {"class": "SpatialExpansion", "spatial_mult": s_enhance},
          {"alpha": 0.2, "class": "LeakyReLU"}]
    if exo:
        hl += [{"class": "Sup3rConcat", "name": exo}, *_conv_block(2, filters),
               {"class": "SkipConnection", "name": "big_skip_2"},
               _res_blocks(2, half, "small_skip_2", filters),
               {"class": "SkipConnection", "name": "big_skip_2"}]
    hl += _conv_block(2, n_out, act=False)
    return hl


def sup3rcc_temporal_d2t_generator(n_out, t_enhance=24, t_roll=12, n_blocks=16, filters=64):
    """sup3r/configs/sup3rcc/gen_trh_1x_24x_2f: depth_to_time head with a temporal roll."""
    return [*_conv_block(3, filters), _res_blocks(3, n_blocks, "small_skip", filters),
            *_conv_block(3, filters), *_conv_block(3, 32 * t_enhance, act=False),
            {"class": "SpatioTemporalExpansion", "temporal_mult": t_enhance,
             "temporal_method": "depth_to_time", "t_roll": t_roll},
            {"alpha": 0.2, "class": "LeakyReLU"}, *_conv_block(3, n_out, act=False)]
