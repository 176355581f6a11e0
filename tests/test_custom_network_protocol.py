"""The object protocol ``sup3r.models`` consumes from ``phygnn.CustomNetwork`` (call
sites sup3r/models/abstract.py:96-101, 319, 1081-1084, base.py:146-150, 235, interface.py:69, 84,
105-121, 384-405, models/utilities.py:23-27, tests/training/test_load_configs.py:65-126), held on
the host (no device needed): construction from ``hidden_layers``, ``name``, ``layers``,
iteration, ``weights`` with keras-style names, layer attributes (``rank``, ``_spatial_mult``,
``_temporal_mult``, ``name``, ``features`` / ``exo_features``), layer groups for ``isinstance``,
seeding, ``save`` / ``load``, exception types."""
import os

import numpy as np
import pytest

from sup3r_b200 import configs as C
from sup3r_b200.network import (SUP3R_EXO_LAYERS, SUP3R_LAYERS, SUP3R_OBS_LAYERS, CustomNetwork)


def _net(hl, seed=3, in_shape=None, exo=None, name="generator"):
    CustomNetwork.seed(seed)
    net = CustomNetwork(hidden_layers=hl, name=name, device="cpu")
    if in_shape is not None:
        net.build(in_shape, exo)
    return net


def test_construction_layers_and_attributes():
    hl = C.sup3rcc_spatial_generator(2, 5, 2, exo="topography", filters=16)
    net = _net(hl)
    assert net.name == "generator"
    assert isinstance(net.layers, list) and len(net) == len(net.layers) == len(list(net))
    assert all(a is b for a, b in zip(net, net.layers))
    # interface.py:84: input rank from layer 0 (every config starts with FlexiblePadding)
    assert type(net.layers[0]).__name__ == "FlexiblePadding" and net.layers[0].rank == 4
    # interface.py:105-121: enhancement factors from the expansion layers' attributes
    mults = [(getattr(l, "_spatial_mult", 1), getattr(l, "_temporal_mult", 1)) for l in net]
    assert int(np.prod([m[0] for m in mults])) == 5 and int(np.prod([m[1] for m in mults])) == 1
    st = _net(C.spatiotemporal_generator(2, 3, (2, 2)))
    assert st.layers[0].rank == 5
    assert int(np.prod([getattr(l, "_spatial_mult", 1) for l in st])) == 3
    assert int(np.prod([getattr(l, "_temporal_mult", 1) for l in st])) == 4
    # models/utilities.py:23-27: layer groups; interface.py:384-405: exo layers by name
    exo_layers = [l for l in net if isinstance(l, SUP3R_LAYERS)]
    assert [l.name for l in exo_layers] == ["topography"]
    assert isinstance(exo_layers[0], SUP3R_EXO_LAYERS)
    assert not isinstance(exo_layers[0], SUP3R_OBS_LAYERS)
    assert set(SUP3R_LAYERS) == set(SUP3R_EXO_LAYERS) | set(SUP3R_OBS_LAYERS)
    obs = _net([{"class": "Sup3rConcatObs", "name": "u_10m_obs", "fill_index": 0},
                {"class": "Sup3rObsModel", "name": "obs", "features": ["u_10m_obs"],
                 "exo_features": ["topography"],
                 "hidden_layers": [{"class": "Conv2D", "filters": 2, "kernel_size": 3,
                                    "padding": "same"}]}])
    assert all(isinstance(l, SUP3R_OBS_LAYERS) for l in obs)
    assert obs.layers[1].features == ["u_10m_obs"] and obs.layers[1].exo_features == ["topography"]


def test_weights_seed_save_load(tmp_path):
    hl = C.sup3rcc_spatial_generator(2, 5, 2, exo="topography", filters=16)
    net = _net(hl, in_shape=(1, 6, 6, 2), exo={"topography": 1})
    assert len(net.weights) == 18
    assert net.weights[0].name == "generator/conv2d/kernel:0"
    assert net.weights[1].name == "generator/conv2d/bias:0"
    assert net.weights[0].numpy().shape == (3, 3, 2, 16)            # keras (k, k, cin, cout)
    assert not net.weights[1].numpy().any()                         # zero biases
    k = net.weights[0].numpy()
    limit = np.sqrt(6.0 / (9 * 2 + 9 * 16))                         # glorot uniform
    assert np.abs(k).max() <= limit and np.abs(k).max() > 0.5 * limit
    same = _net(hl, in_shape=(1, 6, 6, 2), exo={"topography": 1})
    other = _net(hl, seed=4, in_shape=(1, 6, 6, 2), exo={"topography": 1})
    assert all(np.array_equal(a, b) for a, b in zip(net.get_weights(), same.get_weights()))
    assert any(not np.array_equal(a, b) for a, b in zip(net.get_weights(), other.get_weights()))
    fp = str(tmp_path / "model_gen.pkl")
    net.save(fp)
    assert os.path.exists(fp)
    loaded = CustomNetwork.load(fp)
    assert loaded.name == "generator" and len(loaded) == len(net)
    assert [w.name for w in loaded.weights] == [w.name for w in net.weights]
    assert all(np.array_equal(a, b) for a, b in zip(net.get_weights(), loaded.get_weights()))


def test_bad_configs_raise():
    with pytest.raises(KeyError):
        CustomNetwork(hidden_layers=[{"class": "NoSuchLayer"}], device="cpu")
    with pytest.raises(TypeError):
        CustomNetwork(hidden_layers=["Conv2D"], device="cpu")


def test_optimizer_protocol():
    """What sup3r.models consumes from a keras optimiser (abstract.py:321-350, 543-587, base.py:
    326-348, tests/training/test_train_gan.py:370-386): ``get_config`` with ``name`` and
    ``learning_rate``, ``from_config``, ``variables`` with ``.name`` / ``.numpy()``,
    ``learning_rate`` comparable to a float, class lookup by config (``ValueError`` otherwise)."""
    import torch
    from sup3r_b200.models.abstract import AbstractSingleModel
    from sup3r_b200.network import Variable
    from sup3r_b200.optimizers import Adam, get_optimizer_class
    opt = AbstractSingleModel.init_optimizer(None, 1e-4)
    assert isinstance(opt, Adam) and opt.learning_rate == 1e-4
    conf = opt.get_config()
    assert conf["name"] == "Adam" and conf["learning_rate"] == 1e-4
    assert (conf["beta_1"], conf["beta_2"], conf["epsilon"]) == (0.9, 0.999, 1e-7)   # keras
    conf.update(learning_rate=0.25)
    new = get_optimizer_class(conf).from_config(conf)
    assert isinstance(new, Adam) and new.learning_rate == 0.25 and new.beta_2 == 0.999
    from_dict = AbstractSingleModel.init_optimizer(
        {"name": "Adam", "beta_1": 0.5, "learning_rate": 3e-4, "not_a_parameter": 1}, 7.0)
    assert from_dict.beta_1 == 0.5 and from_dict.learning_rate == 3e-4
    assert AbstractSingleModel.init_optimizer(opt, 7.0) is opt
    w = Variable("generator/conv2d/kernel:0", torch.zeros(2, 3))
    opt.slots_for(w)
    names = [v.name for v in opt.variables]
    assert names == ["Adam/iteration:0", "Adam/m/generator/conv2d/kernel:0",
                     "Adam/v/generator/conv2d/kernel:0"]
    assert opt.variables[0].numpy() == 0 and opt.variables[1].numpy().shape == (2, 3)
    state = AbstractSingleModel.get_optimizer_state(opt)
    assert state["learning_rate"] == 1e-4
    with pytest.raises(ValueError):
        get_optimizer_class({"name": "NoSuchOptimizer"})
