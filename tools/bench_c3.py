"""BASELINE configs[2]: sup3rcc wind 2-step chain on one 20x20x72 LR chunk (timing + kernel mix):
step 1 = gen_wind_1x_24x_6f (temporal 24x, 5-D), step 2 = gen_wind_5x_1x_6f (spatial 5x, 4-D, with
topography Sup3rConcat)."""
import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from sup3r_b200.models import Sup3rGan
from sup3r_b200 import configs as C
REF = os.path.join(ROOT, "sup3r_b200", "configs")
def load(name):
    for sub in ("sup3rcc", "spatiotemporal", "spatial", ""):
        fp = os.path.join(REF, sub, name)
        if os.path.exists(fp):
            return json.load(open(fp))["hidden_layers"]
    return None
g1 = load("gen_wind_1x_24x_6f.json") or C.sup3rcc_temporal_d2t_generator(6, 24, 12)
g2 = C.sup3rcc_spatial_generator(6, 5, 16, exo="topography")
feats = [f"f{i}" for i in range(6)]
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
Sup3rGan.seed(0)
m1 = Sup3rGan(g1, C.discriminator(3, "same", (8,)), precision=prec,
              meta={"lr_features": feats, "hr_out_features": feats, "s_enhance": 1, "t_enhance": 24})
m2 = Sup3rGan(g2, C.discriminator(2, "same", (8,)), precision=prec,
              meta={"lr_features": feats, "hr_out_features": feats, "hr_exo_features": ["topography"],
                    "s_enhance": 5, "t_enhance": 1})
rng = np.random.default_rng(0)
x = rng.standard_normal((1, 20, 20, 72, 6)).astype(np.float32)
topo = rng.standard_normal((100, 100, 1)).astype(np.float32)
def chain():
    y1 = m1.generate(x)                                   # (1, 20, 20, 1728, 6)
    x2 = np.transpose(y1[0], (2, 0, 1, 3))                # (1728, 20, 20, 6)
    exo = {"topography": {"steps": [{"model": 0, "combine_type": "layer", "data": topo}]}}
    outs = []
    for i in range(0, x2.shape[0], 432):
        outs.append(m2.generate(x2[i:i + 432], exogenous_data=exo))
    return np.concatenate(outs, 0)
from sup3r_b200.models import MultiStepGan
ms = MultiStepGan([m1, m2])
exo_ms = {"topography": {"steps": [{"model": 1, "combine_type": "layer", "data": topo}]}}
y = chain()
print("out", y.shape)
y_ms = ms.generate(x, exogenous_data=exo_ms)
print("MultiStepGan out", y_ms.shape, "max |diff| vs manual chain", float(np.abs(y_ms - y).max()))
torch.cuda.synchronize(); t0 = time.perf_counter(); y_ms = ms.generate(x, exogenous_data=exo_ms); torch.cuda.synchronize()
print(f"MultiStepGan.generate (device-resident intermediates): {(time.perf_counter()-t0)*1e3:.1f} ms wall")
torch.cuda.synchronize(); t0 = time.perf_counter(); y = chain(); torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{prec}: 2-step chain on one 20x20x72 chunk: {dt*1e3:.1f} ms wall -> {20*20*72/dt/1e3:.1f} k LR voxels/s")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    chain(); torch.cuda.synchronize()
agg = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        d = agg.setdefault(e.name[:60], [0, 0.0]); d[0] += 1; d[1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"device time {tot/1e3:.1f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"  {k:60s} n={v[0]:4d} {v[1]/1e3:8.2f} ms")

# ---- the tiler on a domain of several such chunks (streamed driver: device-side crop + output
# check, pinned D2H on a copy stream overlapped with the next chunk's generator pass)
from sup3r_b200.pipeline import ArrayInputHandler, ForwardPass, ForwardPassStrategy
dom = rng.standard_normal((40, 40, 72, 6)).astype(np.float32)
topo_dom = rng.standard_normal((200, 200, 1)).astype(np.float32)
def exo_dom():
    return {"topography": {"steps": [{"model": 1, "combine_type": "layer", "data": topo_dom.copy(),
                                      "s_enhance": 5, "t_enhance": 24}]}}
for workers in (1, 2):
    dts = []
    for _ in range(2):
        strat = ForwardPassStrategy(model=ms, input_handler=ArrayInputHandler(dom, feats),
                                    fwp_chunk_shape=(20, 20, 72), spatial_pad=0, temporal_pad=0,
                                    pass_workers=workers, exo_data=exo_dom())
        torch.cuda.synchronize(); t0 = time.perf_counter()
        outs = ForwardPass.run(strat, 0)
        torch.cuda.synchronize(); dts.append(time.perf_counter() - t0)
    n = strat.n_chunks
    print(f"ForwardPass.run, {n} chunks of 20x20x72 (+ topography), pass_workers={workers} "
          f"({'serial reference-shaped driver' if workers == 1 else 'streamed driver'}): "
          f"{min(dts) / n * 1e3:.1f} ms wall per chunk -> {dom[..., 0].size / min(dts) / 1e3:.1f} k LR voxels/s")
    del outs
