"""Why the backward pass runs on a scaled loss: generator weight gradients of one training step at
BASELINE configs[3] shapes, tensor-core path (fp16c operands) with and without the power-of-two
loss scale, against the fp32 kernels.   python tools/grad_scale_check.py [batch]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sup3r_b200.models import Sup3rGan
from sup3r_b200 import configs as C

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
feats = [f"f{i}" for i in range(6)]
Sup3rGan.seed(0)
m = Sup3rGan(C.spatiotemporal_generator(6, 2, (2, 2, 3)), C.discriminator(3, "same", (1024,)),
             loss="MeanAbsoluteError",
             meta={"lr_features": feats, "hr_out_features": feats, "s_enhance": 2, "t_enhance": 12})
rng = np.random.default_rng(0)
lr = rng.standard_normal((B, 16, 16, 4, 6)).astype(np.float32)
hr = rng.standard_normal((B, 32, 32, 48, 6)).astype(np.float32)
m.init_weights(lr.shape, hr.shape)
kw = dict(weight_gen_advers=1e-3, train_gen=True, train_disc=False)


def grads(precision, scaled):
    m.precision = precision
    if not scaled:
        m.grad_loss_scale = lambda hi, tg=True: 1.0
    elif "grad_loss_scale" in m.__dict__:
        del m.__dict__["grad_loss_scale"]
    g, det = m.get_single_grad(lr, hr, m.generator_weights, **kw)
    return [t.detach().cpu().numpy().astype(np.float64) for t in g], float(det["loss_gen"])


ref, l32 = grads("fp32", False)
for scaled in (False, True):
    got, l16 = grads("fp16c", scaled)
    errs = [np.abs(a - b).max() / max(np.abs(b).max(), 1e-30) for a, b in zip(got, ref)]
    print(f"fp16c tensor-core backward, loss scale {'on ' if scaled else 'off'}: worst weight-grad "
          f"rel err {max(errs):.3e}, median {np.median(errs):.3e} (loss {l16:.6f} vs fp32 {l32:.6f}; "
          f"scale {m.grad_loss_scale(torch.empty(hr.shape)):.0f})")
print("max |grad| of the first / last generator kernel (fp32):", np.abs(ref[0]).max(), np.abs(ref[-2]).max())
