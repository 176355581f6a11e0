"""Relative error (max |diff| / max |ref|) of the bf16 / bf16x3 modes against the fp32 kernels on
the north-star generator, one 16x16x24x4 chunk."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from sup3r_b200.models import Sup3rGan
from sup3r_b200 import configs as C
Sup3rGan.seed(0)
m = Sup3rGan(bench.gen_config(), C.discriminator(3, "same", (2048, 1024)))
x = np.random.default_rng(1).standard_normal((1, *bench.LR_CHUNK)).astype(np.float32)
ref = m.generate(x, precision="fp32")
for p in ("bf16x3", "bf16"):
    y = m.generate(x, precision=p)
    print(p, "max rel err vs fp32 kernels:", float(np.abs(y - ref).max() / np.abs(ref).max()),
          " rms rel:", float(np.sqrt(np.mean((y - ref) ** 2)) / np.sqrt(np.mean(ref ** 2))))
