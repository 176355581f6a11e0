"""Bridge for genuine TensorFlow parity (run this in a REAL sup3r environment, not here):
exports a sup3r model's weights + a reference output to neutral files that ``sup3r_b200`` can
load without TensorFlow.

    # a trained model
    python export_phygnn_weights.py /path/to/sup3r_model_dir out_dir
    # no trained model at hand: a randomly initialised one from generator / discriminator configs
    # (e.g. sup3r/configs/spatiotemporal/gen_3x_4x_2f.json + disc.json); biases get small random
    # values so that the bias path is pinned too
    python export_phygnn_weights.py --config gen.json disc.json out_dir

Writes out_dir/{gen,disc}_weights.npz (arrays w000, w001, ... in keras order: kernel, bias per
layer), {gen,disc}_weight_names.json, {gen,disc}_hidden_layers.json (the phygnn layer configs),
copies model_params.json (trained models), and stores a seeded low-res input with the reference
generator's output (golden.npz, ``norm_in = un_norm_out = False``) so that ``tests`` can pin the
kernels against real TF numbers: drop out_dir under ``tests/golden/phygnn/<name>/`` --
``tests/test_phygnn_golden.py`` picks it up (``sup3r_b200.interop.load_exported_model`` is the
loader).  Models with exogenous / observation layers are refused (the slot feeds no exo data).
"""
import json
import os
import shutil
import sys

import numpy as np


def _hidden_layers(net, config_file=None):
    if config_file is not None:
        with open(config_file) as f:
            cfg = json.load(f)
        return cfg["hidden_layers"] if isinstance(cfg, dict) else cfg
    for get in (lambda: net.model_params["hidden_layers"], lambda: net.hidden_layers,
                lambda: net._hidden_layers_kwargs):
        try:
            return get()
        except Exception:      # noqa: BLE001
            continue
    raise RuntimeError("could not read the hidden_layers config of the network")


def _export(model, out_dir, configs=(None, None)):
    if list(getattr(model, "hr_exo_features", [])) or list(getattr(model, "obs_features", [])):
        raise SystemExit("this model has exogenous / observation layers: export one without "
                         "them for the parity slot")
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(42)
    n_in = len(model.lr_features) or 2
    shape = (1, 10, 10, 6, n_in) if model.is_5d else (4, 10, 10, n_in)
    x = rng.standard_normal(shape).astype(np.float32)
    if configs[0] is not None:
        model.generate(x, norm_in=False, un_norm_out=False)     # builds the weights
        for w in model.generator.weights:
            if "bias" in w.name:
                w.assign((rng.standard_normal(tuple(w.shape)) * 0.05).astype(np.float32))
    y = np.asarray(model.generate(x, norm_in=False, un_norm_out=False))
    nets = [("gen", model.generator, configs[0])]
    if len(model.discriminator.weights) > 0:
        nets.append(("disc", model.discriminator, configs[1]))
    for tag, net, cfg in nets:
        arrs = {f"w{i:03d}": w.numpy() for i, w in enumerate(net.weights)}
        np.savez_compressed(os.path.join(out_dir, f"{tag}_weights.npz"), **arrs)
        with open(os.path.join(out_dir, f"{tag}_weight_names.json"), "w") as f:
            json.dump([w.name for w in net.weights], f)
        with open(os.path.join(out_dir, f"{tag}_hidden_layers.json"), "w") as f:
            json.dump({"hidden_layers": _hidden_layers(net, cfg)}, f)
    np.savez_compressed(os.path.join(out_dir, "golden.npz"), low_res=x, hi_res=y)
    print("wrote", out_dir, "input", x.shape, "output", y.shape)


def main(argv):
    from sup3r.models import Sup3rGan  # needs tensorflow + phygnn
    if argv and argv[0] == "--config":
        fp_gen, fp_disc, out_dir = argv[1:4]
        Sup3rGan.seed(0)
        _export(Sup3rGan(fp_gen, fp_disc), out_dir, configs=(fp_gen, fp_disc))
        return
    model_dir, out_dir = argv[:2]
    _export(Sup3rGan.load(model_dir), out_dir)
    shutil.copy(os.path.join(model_dir, "model_params.json"), out_dir)


if __name__ == "__main__":
    main(sys.argv[1:])
