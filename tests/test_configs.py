"""Config builders reproduce the reference's JSON network definitions exactly, and every
reference config maps shapes like tests/training/test_load_configs.py:17-133 says (shape
propagation only; no GPU).  Skipped where /root/reference is absent (the GPU box)."""
import glob
import json
import os

import pytest

from sup3r_b200 import configs as C
from sup3r_b200.network import CustomNetwork, expand_hidden_layers

REF = "/root/reference/sup3r/configs"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def ref(path):
    return json.load(open(os.path.join(REF, path)))["hidden_layers"]


@needs_ref
def test_builders_equal_reference_files():
    assert C.spatiotemporal_generator(14, 2, (2, 2, 3)) == ref("spatiotemporal/gen_2x_12x_14f.json")
    assert C.spatiotemporal_generator(2, 3, (2, 2)) == ref("spatiotemporal/gen_3x_4x_2f.json")
    assert C.spatiotemporal_generator(1, 3, (2, 2)) == ref("spatiotemporal/gen_3x_4x_1f.json")
    assert C.spatiotemporal_generator(10, 3, (2, 2)) == ref("spatiotemporal/gen_3x_4x_10f.json")
    assert C.spatiotemporal_generator(14, 3, (2, 2), head_filters=576) \
        == ref("spatiotemporal/gen_3x_4x_14f.json")
    assert C.spatiotemporal_generator(2, 2, (2,)) == ref("spatiotemporal/gen_2x_2x_2f.json")
    assert C.spatiotemporal_generator(3, 4, (2, 2, 2, 3)) == ref("spatiotemporal/gen_4x_24x_3f.json")
    assert C.spatial_generator(2, (2,)) == ref("spatial/gen_2x_2f.json")
    assert C.spatial_generator(1, (2,)) == ref("spatial/gen_2x_1f.json")
    assert C.spatial_generator(2, (2, 5)) == ref("spatial/gen_10x_2f.json")
    assert C.discriminator(2, "valid", (1024,)) == ref("spatial/disc.json")
    assert C.discriminator(3, "valid", (2048, 1024)) == ref("spatiotemporal/disc.json")
    assert C.sup3rcc_spatial_generator(6, 5, 16, exo="topography") \
        == ref("sup3rcc/gen_wind_5x_1x_6f.json")
    assert C.sup3rcc_spatial_generator(1, 5, 16) == ref("sup3rcc/gen_solar_5x_1x_1f.json")
    assert C.sup3rcc_temporal_d2t_generator(2, 24, 12) == ref("sup3rcc/gen_trh_1x_24x_2f.json")
    tdisc = "/root/reference/tests/data/config_disc_st_test.json"
    assert C.discriminator(3, "same", (2048, 1024)) == json.load(open(tdisc))["hidden_layers"]


@needs_ref
@pytest.mark.parametrize("fp", sorted(glob.glob(os.path.join(REF, "spatiotemporal", "gen_*.json"))))
def test_reference_st_gen_configs_shapes(fp):
    """(test_load_configs.py:17-67): ones of shape (n, 7, 7, 4, n_in) -> enhanced shape."""
    name = os.path.basename(fp).replace(".json", "")
    _, s, t, f = name.split("_")
    s, t, f = int(s[:-1]), int(t[:-1]), int(f[:-1])
    net = CustomNetwork(json.load(open(fp))["hidden_layers"], name="generator")
    for n_in in (f, f + 2):
        assert net.output_shape((3, 7, 7, 4, n_in)) == (3, 7 * s, 7 * s, 4 * t, f)


@needs_ref
@pytest.mark.parametrize("fp", sorted(glob.glob(os.path.join(REF, "spatial", "gen_*.json"))))
def test_reference_s_gen_configs_shapes(fp):
    name = os.path.basename(fp).replace(".json", "")
    _, s, f = name.split("_")
    s, f = int(s[:-1]), int(f[:-1])
    net = CustomNetwork(json.load(open(fp))["hidden_layers"], name="generator")
    assert net.output_shape((4, 10, 10, f)) == (4, 10 * s, 10 * s, f)


@needs_ref
def test_reference_disc_and_sup3rcc_configs_shapes():
    net = CustomNetwork(ref("spatiotemporal/disc.json"), name="discriminator")
    assert net.output_shape((2, 62, 62, 62, 3)) == (2, 1)
    with pytest.raises(RuntimeError):
        net.output_shape((2, 20, 20, 20, 3))  # valid-padded disc needs >= 61 points per dim
    net = CustomNetwork(ref("spatial/disc.json"), name="discriminator")
    assert net.output_shape((2, 64, 64, 2)) == (2, 1)
    expect = {"gen_solar_1x_8x_1f": ((1, 8, 8, 6, 1), (1, 8, 8, 48, 1)),
              "gen_solar_5x_1x_1f": ((3, 8, 8, 1), (3, 40, 40, 1)),
              "gen_trh_1x_24x_2f": ((1, 8, 8, 5, 2), (1, 8, 8, 120, 2)),
              "gen_wind_1x_24x_6f": ((1, 8, 8, 5, 6), (1, 8, 8, 120, 6)),
              "gen_wind_3x_4x_2f": ((1, 8, 8, 5, 2), (1, 24, 24, 20, 2)),
              "gen_wind_5x_1x_6f": ((3, 8, 8, 6), (3, 40, 40, 6))}
    for name, (i, o) in expect.items():
        net = CustomNetwork(ref(f"sup3rcc/{name}.json"), name="generator")
        assert net.output_shape(i) == o, name


def test_expand_repeat_and_shared_skips():
    hl = C.spatiotemporal_generator(4, 5, (2, 2, 3), head_filters=200)
    flat = expand_hidden_layers(hl)
    assert len(flat) == 172
    net = CustomNetwork(hl, name="generator")
    skips = [lyr for lyr in net.layers if type(lyr).__name__ == "SkipConnection"]
    assert len({id(s) for s in skips}) == 2 and len(skips) == 34
    assert net.layers[0].rank == 5
    assert net.output_shape((1, 16, 16, 24, 4)) == (1, 80, 80, 288, 4)
    with pytest.raises(KeyError):
        CustomNetwork([{"class": "NotALayer"}])
    with pytest.raises(RuntimeError):
        net.output_shape((1, 3, 16, 24, 4))  # REFLECT pad 3 needs > 3 points... conv too small


def test_derived_config_files_match_builders():
    here = os.path.join(os.path.dirname(C.__file__), "spatiotemporal")
    got = json.load(open(os.path.join(here, "gen_5x_12x_4f.json")))["hidden_layers"]
    assert got == C.spatiotemporal_generator(4, 5, (2, 2, 3), head_filters=200)
    got = json.load(open(os.path.join(here, "gen_5x_24x_4f.json")))["hidden_layers"]
    assert got == C.spatiotemporal_generator(4, 5, (2, 2, 2, 3), head_filters=200)
    got = json.load(open(os.path.join(here, "disc_same.json")))["hidden_layers"]
    assert got == C.discriminator(3, "same", (2048, 1024))
