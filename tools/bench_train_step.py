"""Wall/device time of one Sup3rGan training step (generator + discriminator gradient steps) on the
BASELINE config[3]-style shapes: generator = gen_2x_12x pattern with 6 in / 6 out features,
LR (B, 16, 16, 4, 6) -> HR (B, 32, 32, 48, 6), 'same'-padded ST discriminator.
  python tools/bench_train_step.py [batch] [steps]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sup3r_b200.models import Sup3rGan
from sup3r_b200 import configs as C
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
K = int(sys.argv[2]) if len(sys.argv) > 2 else 12
feats = [f"f{i}" for i in range(6)]
Sup3rGan.seed(0)
m = Sup3rGan(C.spatiotemporal_generator(6, 2, (2, 2, 3)), C.discriminator(3, "same", (1024,)),
             learning_rate=1e-4, loss="MeanAbsoluteError",
             meta={"lr_features": feats, "hr_out_features": feats, "s_enhance": 2, "t_enhance": 12})
rng = np.random.default_rng(0)
lr = rng.standard_normal((B, 16, 16, 4, 6)).astype(np.float32)
hr = rng.standard_normal((B, 32, 32, 48, 6)).astype(np.float32)
m.generator.build(lr.shape)
m.discriminator.build(hr.shape)
m.init_weights(lr.shape, hr.shape) if hasattr(m, "init_weights") else None
dev = m.torch_device()
lr_t, hr_t = torch.tensor(lr, device=dev), torch.tensor(hr, device=dev)
times = []
for i in range(K + 1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d1 = m.run_gradient_descent(lr_t, hr_t, m.generator_weights, weight_gen_advers=1e-3,
                                train_gen=True, train_disc=False)
    d2 = m.run_gradient_descent(lr_t, hr_t, m.discriminator_weights, weight_gen_advers=1e-3,
                                train_gen=False, train_disc=True)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
# (steps 0-1 run eagerly, step 2 captures the CUDA graphs of the two gradient steps)
print(f"batch {B}: gen + disc gradient step {np.median(times[4:])*1e3:.1f} ms (median of steps 4+; "
      f"steps 0-3: {' '.join(f'{t*1e3:.0f}' for t in times[:4])} ms; "
      f"graph {dict(m._graphed_steps.stats)}); loss_gen {float(d1['loss_gen']):.4f} "
      f"loss_disc {float(d2['loss_disc']):.4f}")
