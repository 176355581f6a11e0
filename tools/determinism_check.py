"""Run-to-run determinism of the gradient step: the same shard gradients computed twice from the
same state (no optimiser step in between) must agree bit for bit.
  python tools/determinism_check.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_distributed_gpu as T

m = T._model()
lr, hr = T._batches()[0]
for tag, weights, kw in (("gen", m.generator_weights, dict(train_gen=True, train_disc=False, compute_disc=True)),
                         ("disc", m.discriminator_weights, dict(train_gen=False, train_disc=True))):
    runs = []
    for _ in range(3):
        g, d = m.get_single_grad(lr, hr, weights, weight_gen_advers=1e-2, **kw)
        runs.append([t.clone().cpu().numpy() for t in g])
    for v, a, b, c in zip(weights, *runs):
        d1 = max(np.abs(a - b).max(), np.abs(a - c).max())
        print(f"{tag} {v.name:40s} max|g| {np.abs(a).max():.3e}  run-to-run diff {d1:.3e}")
