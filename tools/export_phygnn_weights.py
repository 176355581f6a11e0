"""Bridge for genuine TensorFlow parity (run this in a REAL sup3r environment, not here):
exports a trained sup3r model's weights + a reference output to neutral files that
``sup3r_b200`` can load without TensorFlow.

    python export_phygnn_weights.py /path/to/sup3r_model_dir out_dir

Writes out_dir/{gen,disc}_weights.npz (arrays w000, w001, ... in keras order: kernel, bias per
layer), {gen,disc}_hidden_layers.json (the phygnn layer configs), copies model_params.json, and
stores a seeded low-res input with the reference generator's output (golden.npz) so that
``tests`` can pin the kernels against real TF numbers: drop out_dir under
``tests/golden/phygnn/<name>/`` -- ``tests/test_phygnn_golden.py`` picks it up
(``sup3r_b200.interop.load_exported_model`` is the loader).
"""
import json
import os
import shutil
import sys

import numpy as np


def main(model_dir, out_dir):
    from sup3r.models import Sup3rGan  # needs tensorflow + phygnn
    os.makedirs(out_dir, exist_ok=True)
    model = Sup3rGan.load(model_dir)
    for tag, net in (("gen", model.generator), ("disc", model.discriminator)):
        arrs = {f"w{i:03d}": w.numpy() for i, w in enumerate(net.weights)}
        np.savez_compressed(os.path.join(out_dir, f"{tag}_weights.npz"), **arrs)
        with open(os.path.join(out_dir, f"{tag}_weight_names.json"), "w") as f:
            json.dump([w.name for w in net.weights], f)
        hl = None
        for get in (lambda: net.model_params["hidden_layers"], lambda: net.hidden_layers,
                    lambda: net._hidden_layers_kwargs):
            try:
                hl = get()
                break
            except Exception:
                continue
        with open(os.path.join(out_dir, f"{tag}_hidden_layers.json"), "w") as f:
            json.dump({"hidden_layers": hl}, f)
    shutil.copy(os.path.join(model_dir, "model_params.json"), out_dir)
    rng = np.random.default_rng(42)
    n_in = len(model.lr_features) or 2
    shape = (1, 10, 10, 6, n_in) if model.is_5d else (4, 10, 10, n_in)
    x = rng.standard_normal(shape).astype(np.float32)
    y = model.generate(x, norm_in=False, un_norm_out=False)
    np.savez_compressed(os.path.join(out_dir, "golden.npz"), low_res=x, hi_res=y)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
