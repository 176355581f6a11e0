"""Data-parallel training step under torchrun (one rank per GPU): batch 4 per GPU of the BASELINE
configs[3] shapes, gradients SUM-all-reduced (sup3r/models/abstract.py:785-841 semantics); device
time per generator + discriminator step, max over ranks.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train_multi.py [steps]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from sup3r_b200 import parallel
from sup3r_b200.models import Sup3rGan
from sup3r_b200 import configs as C
K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
parallel.init_from_env("nccl")
feats = [f"f{i}" for i in range(6)]
Sup3rGan.seed(0)
m = Sup3rGan(C.spatiotemporal_generator(6, 2, (2, 2, 3)), C.discriminator(3, "same", (1024,)),
             learning_rate=1e-4, loss="MeanAbsoluteError", default_device=f"/gpu:{local}",
             meta={"lr_features": feats, "hr_out_features": feats, "s_enhance": 2, "t_enhance": 12})
B = 4 * world                      # the global batch every rank is fed (tf.split semantics)
rng = np.random.default_rng(0)
lr = rng.standard_normal((B, 16, 16, 4, 6)).astype(np.float32)
hr = rng.standard_normal((B, 32, 32, 48, 6)).astype(np.float32)
m.init_weights((4, 16, 16, 4, 6), (4, 32, 32, 48, 6))
parallel.broadcast_weights([m.generator, m.discriminator])
dev = m.torch_device()
lr_t, hr_t = torch.tensor(lr, device=dev), torch.tensor(hr, device=dev)
ts = []
W = 4    # untimed: 2 eager steps, the CUDA-graph capture, 1 replay
for i in range(K + W):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m.run_gradient_descent(lr_t, hr_t, m.generator_weights, weight_gen_advers=1e-3, train_gen=True,
                           train_disc=False, multi_gpu=world > 1)
    d = m.run_gradient_descent(lr_t, hr_t, m.discriminator_weights, weight_gen_advers=1e-3,
                               train_gen=False, train_disc=True, multi_gpu=world > 1)
    e1.record()
    torch.cuda.synchronize()
    if i >= W:
        ts.append(e0.elapsed_time(e1))
t = torch.tensor([float(np.mean(ts))], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if parallel.rank() == 0:
    print(f"{world} GPU(s): global batch {B} ({B // world} per GPU): gen + disc step {float(t):.1f} ms "
          f"(max over ranks) -> {B / float(t) * 1e3:.1f} samples/s; loss_disc {float(d['loss_disc']):.4f}")
if world > 1:
    dist.destroy_process_group()
