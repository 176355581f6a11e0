"""Per-kernel device time of one Sup3rGan training step (generator step + discriminator step) on
the BASELINE configs[3]-style shapes (see tools/bench_train_step.py) via torch.profiler.
  python tools/train_breakdown.py [batch]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from sup3r_b200.models import Sup3rGan
from sup3r_b200 import configs as C
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
feats = [f"f{i}" for i in range(6)]
Sup3rGan.seed(0)
m = Sup3rGan(C.spatiotemporal_generator(6, 2, (2, 2, 3)), C.discriminator(3, "same", (1024,)),
             learning_rate=1e-4, loss="MeanAbsoluteError",
             meta={"lr_features": feats, "hr_out_features": feats, "s_enhance": 2, "t_enhance": 12})
rng = np.random.default_rng(0)
lr = rng.standard_normal((B, 16, 16, 4, 6)).astype(np.float32)
hr = rng.standard_normal((B, 32, 32, 48, 6)).astype(np.float32)
m.init_weights(lr.shape, hr.shape)
dev = m.torch_device()
lr_t, hr_t = torch.tensor(lr, device=dev), torch.tensor(hr, device=dev)


def step():
    m.run_gradient_descent(lr_t, hr_t, m.generator_weights, weight_gen_advers=1e-3,
                           train_gen=True, train_disc=False)
    m.run_gradient_descent(lr_t, hr_t, m.discriminator_weights, weight_gen_advers=1e-3,
                           train_gen=False, train_disc=True)


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in evs:
    d = agg.setdefault(e.name[:90], [0, 0.0])
    d[0] += 1
    d[1] += e.device_time
tot = sum(v[1] for v in agg.values())
evs.sort(key=lambda e: e.time_range.start)
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"training step batch {B}: {len(evs)} kernels, device sum {tot / 1e3:.1f} ms, span {span / 1e3:.1f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"{k:90s} n={v[0]:4d} {v[1] / 1e3:8.2f} ms {100 * v[1] / tot:5.1f}%")
# the longest individual launches (which layer shapes carry the step)
print("longest launches:")
for e in sorted(evs, key=lambda e: -e.device_time)[:28]:
    print(f"  {e.device_time:8.1f} us  {e.name[:70]}")
# histogram of the ring kernel's launch durations
ring = sorted(e.device_time for e in evs if "zring" in e.name)
if ring:
    import numpy as _np
    r = _np.array(ring)
    print(f"ring kernel: n={len(r)} median {_np.median(r):.1f} us, <15us: {(r < 15).sum()}, "
          f"15-50us: {((r >= 15) & (r < 50)).sum()}, >=50us: {(r >= 50).sum()} "
          f"(sum {r[r >= 50].sum() / 1e3:.2f} ms)")
