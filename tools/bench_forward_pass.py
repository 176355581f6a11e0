"""Throughput of the public ForwardPass.run path (host chunking + generate + output check) on a
synthetic LR domain cut into 16x16x24 chunks: python tools/bench_forward_pass.py [s] [t] [workers]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from sup3r_b200.models import Sup3rGan
from sup3r_b200 import configs as C
from sup3r_b200.pipeline import ArrayInputHandler, ForwardPass, ForwardPassStrategy
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 96
workers = int(sys.argv[3]) if len(sys.argv) > 3 else 8
Sup3rGan.seed(0)
feats = ["u_10m", "v_10m", "u_100m", "v_100m"]
m = Sup3rGan(bench.gen_config(), C.discriminator(3, "same", (2048, 1024)), precision="bf16",
             meta={"lr_features": feats, "hr_out_features": feats, "s_enhance": 5, "t_enhance": 12})
data = np.random.default_rng(0).standard_normal((S, S, T, 4)).astype(np.float32)
for rep in range(2):
    strat = ForwardPassStrategy(model=m, input_handler=ArrayInputHandler(data, feats),
                                fwp_chunk_shape=(16, 16, 24), spatial_pad=0, temporal_pad=0,
                                pass_workers=workers)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = ForwardPass.run(strat, 0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"ForwardPass.run: {strat.n_chunks} chunks, {S*S*T} LR voxels in {dt*1e3:.1f} ms -> "
          f"{S*S*T/dt/1e6:.2f} M LR voxels/s (pass_workers={workers})")
