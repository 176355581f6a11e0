"""Host-side routing of the fused plan (no GPU needed): which convolution of a reference network
goes to which kernel family.  Guards the fusion (172 layers -> 38 conv steps for the north-star
generator, sup3r/configs/spatiotemporal/gen_*.json dialect) and the tcgen05 eligibility rules."""
import pytest

from sup3r_b200 import configs as C
from sup3r_b200.network import CustomNetwork
from sup3r_b200.plan import FusedConv, _umma_ok, build_steps


def _routes(hl, in_shape, precision):
    net = CustomNetwork(hl, name="generator", device="cpu")
    steps = build_steps(net.layers)
    shp = tuple(in_shape)
    out = []
    for st in steps:
        if isinstance(st, FusedConv):
            out.append((shp[-1], st.conv.filters, st.r, st.m, _umma_ok(st, shp, precision)))
            spec = st.spec(shp)
            n = shp[0]
            dims = (1, *shp[1:-1]) if len(shp) == 4 else shp[1:-1]
            # shape propagation through the conv + scatter without the C library
            r = st.r
            m = st.m
            if len(shp) == 5:
                shp = (n, shp[1] * r, shp[2] * r, shp[3] * m, st.conv.filters // (r * r * (m if st.method == 1 else 1)))
            else:
                shp = (n, shp[1] * r, shp[2] * r, st.conv.filters // (r * r))
            del spec, dims
        else:
            layer = getattr(st, "layer", None)
            if layer is not None:
                shp = layer.out_shape(shp) if hasattr(layer, "out_shape") else shp
    return steps, out


def test_north_star_generator_collapses_to_38_conv_steps():
    hl = C.spatiotemporal_generator(4, 5, (2, 2, 3), head_filters=200)
    steps, routes = _routes(hl, (8, 16, 16, 24, 4), "bf16")
    assert len([s for s in steps if isinstance(s, FusedConv)]) == 38
    assert len(steps) == 38                                   # nothing left un-fused
    # first layer: 4 -> 64 on the tensor cores (zero-padded channels) in bf16 mode only
    assert routes[0][:2] == (4, 64) and routes[0][4]
    assert not _routes(hl, (8, 16, 16, 24, 4), "bf16x3")[1][0][4]
    # body: 64 -> 64; head 64 -> 200 with 5x depth_to_space; output conv 8 -> 4 not on tcgen05
    assert all(r[4] for r in routes[1:-1])
    assert routes[-2][:3] == (64, 200, 5)
    assert routes[-1][:2] == (8, 4) and not routes[-1][4]
    # fp32 mode never uses the tensor-core path
    assert not any(r[4] for r in _routes(hl, (8, 16, 16, 24, 4), "fp32")[1])


def test_wide_scatter_heads_and_concat_conv_routes():
    hl = C.sup3rcc_spatial_generator(6, 5, 16, exo="topography")
    _, routes = _routes(hl, (4, 20, 20, 6), "bf16")
    wide = [r for r in routes if r[1] == 1600]
    assert wide and wide[0][4]                                # 64 -> 1600: channel slices
    concat = [r for r in routes if r[0] == 65]
    assert concat and concat[0][4]                            # 64 + exo channel: split conv
    assert routes[0][:2] == (6, 64) and routes[0][4]          # narrow 2-D input, padded route
    _, routes_c = _routes(hl, (4, 20, 20, 6), "fp16c")        # the benchmarked mode too
    assert [r for r in routes_c if r[0] == 65][0][4]
    hl_t = C.sup3rcc_temporal_d2t_generator(6, 24, 12)
    _, routes_t = _routes(hl_t, (1, 20, 20, 72, 6), "bf16")
    d2t = [r for r in routes_t if r[3] == 24]
    assert d2t and d2t[0][4]                                  # 64 -> 768 with 24x depth_to_time
