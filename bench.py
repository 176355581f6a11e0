"""Benchmark of the sup3r hot path on B200: low-res voxels / second through the generator of
the north-star configuration (derived 5x spatial / 12x temporal / 4 feature spatiotemporal
Sup3rGan, LR chunks 16x16x24 -> HR 80x80x288; BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--chunks B]
                  [--precision fp16c|bf16|bf16x3|fp32]
  python bench.py --impl reference ...     # CPU stand-in for the reference's TensorFlow path

One "step" = one generator pass over a batch of B synthetic LR chunks (default 37: 36.0 work
items of the body kernel per SM; the end-to-end run tiles a domain in batches of 8) (seeded normal fields,
random-init weights of the named architecture).  The default precision ``fp16c`` (fp16 operands
+ e4m3 correction rows, one kind::f16 and one kind::f8f6f4 tcgen05 pass per layer) is the
fastest mode INSIDE the north star's 1e-3 tolerance; the line's ``parity`` block measures it
against the CPU oracle on the same weights and inputs, ``precision_modes`` times the others.
``value`` is device-timed (CUDA events around each step, L2 flushed between steps, max over
ranks) with inputs resident in HBM; ``e2e`` is the same metric through the reference-facing
call ``ForwardPass.run(strategy, node_index=rank)`` on a synthetic LR domain sharded by
``strategy.node_chunks``: host chunking, H2D of every LR batch and D2H of every HR result
(float16; the fp32 figure is reported next to it) inside the timed region.  Under torchrun every rank runs
the same per-GPU work on its own chunks (weak scaling; the only collective is the weight
broadcast at start-up, as the reference's nodes share nothing but the model files).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LR_CHUNK = (16, 16, 24, 4)          # s1, s2, t, features
S_ENH, T_ENH = 5, 12
WORKLOAD = "Sup3rGan spatiotemporal 5x/12x/4f generator forward, LR chunks 16x16x24x4 -> 80x80x288x4"
METRIC = "lr_voxels_per_sec"
E2E_BATCHES = 8                     # batches per rank of the end-to-end ForwardPass.run domain


def gen_config():
    from sup3r_b200 import configs as C
    return C.spatiotemporal_generator(4, S_ENH, (2, 2, 3), head_filters=200)


def algorithmic_flops_per_chunk(hl, lr_shape):
    """2 * sum over convs of (output voxels after crop) * taps * cin * cout."""
    from sup3r_b200.network import CustomNetwork, _Conv
    net = CustomNetwork(hl, name="generator", device="cpu")
    shp = (1, *lr_shape)
    flops = 0.0
    pending = None
    for lyr in net.layers:
        out = lyr.out_shape(shp)
        if isinstance(lyr, _Conv):
            pending = (lyr, shp[-1])
        elif pending is not None and type(lyr).__name__.startswith("Cropping"):
            conv, cin = pending
            vox = np.prod(out[1:-1])
            flops += 2.0 * vox * np.prod(conv.kernel_size) * cin * conv.filters
            pending = None
        shp = out
    return flops


def network_flops(hl, in_shape):
    """2 * sum over conv / dense layers of MACs for one forward pass of ``hl`` on ``in_shape``
    (conv: output voxels after a directly following Cropping * taps * cin * cout)."""
    from sup3r_b200.network import CustomNetwork, Dense, _Conv
    net = CustomNetwork(hl, name="n", device="cpu")
    shp = tuple(in_shape)
    flops, pending = 0.0, None

    def flush(out_shape):
        nonlocal flops, pending
        if pending is not None:
            conv, cin = pending
            flops += 2.0 * np.prod(out_shape[:-1]) * np.prod(conv.kernel_size) * cin * conv.filters
            pending = None
    for lyr in net.layers:
        out = lyr.out_shape(shp)
        if isinstance(lyr, _Conv):
            flush(shp)
            pending = (lyr, shp[-1])
        elif type(lyr).__name__.startswith("Cropping") and pending is not None:
            flush(out)
        else:
            flush(shp)
            if isinstance(lyr, Dense):
                flops += 2.0 * np.prod(shp) * lyr.units
        shp = out
    flush(shp)
    return flops


def train_step_block(dev, peaks, steps=6):
    """BASELINE configs[3]-style training step: generator = gen_2x_12x pattern with 6 in / 6 out
    features, LR (4, 16, 16, 4, 6) -> HR (4, 32, 32, 48, 6), 'same'-padded ST discriminator, Adam
    1e-4, MeanAbsoluteError content loss: one generator gradient step + one discriminator
    gradient step (sup3r/models/base.py:944-1031)."""
    import torch
    from sup3r_b200.models import Sup3rGan
    from sup3r_b200 import configs as C
    B = 4
    feats = [f"f{i}" for i in range(6)]
    gen_hl = C.spatiotemporal_generator(6, 2, (2, 2, 3))
    disc_hl = C.discriminator(3, "same", (1024,))
    Sup3rGan.seed(0)
    m = Sup3rGan(gen_hl, disc_hl, learning_rate=1e-4, loss="MeanAbsoluteError",
                 default_device=f"/gpu:{dev.index or 0}",
                 meta={"lr_features": feats, "hr_out_features": feats, "s_enhance": 2,
                       "t_enhance": 12})
    rng = np.random.default_rng(0)
    lr_shape, hr_shape = (B, 16, 16, 4, 6), (B, 32, 32, 48, 6)
    m.init_weights(lr_shape, hr_shape)
    lr_t = torch.tensor(rng.standard_normal(lr_shape).astype(np.float32), device=dev)
    hr_t = torch.tensor(rng.standard_normal(hr_shape).astype(np.float32), device=dev)

    def step():
        m.run_gradient_descent(lr_t, hr_t, m.generator_weights, weight_gen_advers=1e-3,
                               train_gen=True, train_disc=False)
        m.run_gradient_descent(lr_t, hr_t, m.discriminator_weights, weight_gen_advers=1e-3,
                               train_gen=False, train_disc=True)
    # the same step issued launch by launch (graphs off): 3 warm-up steps, 3 timed
    prev = os.environ.get("SUP3R_B200_TRAIN_GRAPH")
    os.environ["SUP3R_B200_TRAIN_GRAPH"] = "0"
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ms_eager = (time.perf_counter() - t0) * 1e3 / 3
    if prev is None:
        os.environ.pop("SUP3R_B200_TRAIN_GRAPH")
    else:
        os.environ["SUP3R_B200_TRAIN_GRAPH"] = prev
    for _ in range(4):     # untimed: 2 eager steps, the CUDA-graph capture of both steps, 1 replay
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    f_gen, f_disc = network_flops(gen_hl, lr_shape), network_flops(disc_hl, hr_shape)
    # generator step: G fwd + dgrad + wgrad, D fwd on (true, gen), D dgrad through the gen branch;
    # discriminator step: G fwd, D fwd x 2, D dgrad + wgrad x 2
    flops = (3 * f_gen + 3 * f_disc) + (f_gen + 6 * f_disc)
    peak = peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"]
    return {"workload": "Sup3rGan training step (generator step + discriminator step), batch 4, "
                        "LR 16x16x4x6 -> HR 32x32x48x6, gen_2x_12x-pattern generator, same-padded "
                        "ST discriminator (BASELINE configs[3] shapes)",
            "ms_per_step": ms, "ms_per_step_eager_launches": ms_eager,
            "cuda_graph": dict(m._graphed_steps.stats),
            "algorithmic_tflops": flops / (ms / 1e3) / 1e12,
            "frac_of_bf16_peak": flops / (ms / 1e3) / 1e12 / peak,
            "algorithmic_flops_per_step": flops,
            "kernels": "generator AND discriminator convolutions: forward + input gradients (fp16c "
                       "operands, 64-channel groups summed through the f32 residual input, stride 2 "
                       "= sampled stride-1 convolution) + weight gradients (fp16, voxels as the K "
                       "dimension of MN-major operands) on tcgen05; dense layers, losses, Adam, "
                       "layers with extents < 2: fp32 CUDA-core kernels.  Each gradient step is "
                       "replayed as one CUDA graph (sup3r_b200/train_graph.py) + one fused "
                       "whole-network Adam launch; ms_per_step_eager_launches is the same step "
                       "issued launch by launch (graphs off, wall clock, after warm-up)"}


class ClockSampler:
    """Streams nvidia-smi SM clocks / throttle reasons (one line every 50 ms) while the timed
    region runs; only samples taken between start() and finish() are kept."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index=0):
        self.index = index
        self.lines = []
        self.proc = None
        self.reader = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for ln in self.proc.stdout:
                self.lines.append((time.perf_counter(), ln.strip()))
        self.reader = threading.Thread(target=pump, daemon=True)
        self.reader.start()
        # wait for the first sample so that the stream is live when the timed region starts
        t_end = time.perf_counter() + 3.0
        while not self.lines and time.perf_counter() < t_end:
            time.sleep(0.01)
        self.t0 = time.perf_counter()

    def finish(self):
        t1 = time.perf_counter()
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()
        rows = [ln.split(",") for t, ln in self.lines if self.t0 <= t <= t1 + 0.06 and ln]
        rows = [[v.strip() for v in r] for r in rows if len(r) >= 6]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        pw = [float(r[6]) for r in rows if len(r) > 6 and r[6].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(rows), "power_w_max": max(pw) if pw else None}


def load_peaks():
    fp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(fp):
        p = json.load(open(fp))
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def cpu_reference(steps, warmup, sample_chunks=1, weights=None, x=None, want_output=False):
    """torch-CPU port of the literal reference op order on all host cores.  ``weights`` / ``x``:
    run on the GPU arm's weights and input chunk (its output is then the parity oracle)."""
    import torch
    from oracle.torch_ref import TorchRefNet
    from oracle import layers_ref as L
    hl = gen_config()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if weights is None:
        layers = L.build_layers(hl)
        L.build_weights(layers, (1, *LR_CHUNK), seed=0)
        weights = L.get_weights(layers)
    net = TorchRefNet(hl, weights, torch.float32)
    if x is None:
        x = np.random.default_rng(42).standard_normal((sample_chunks, *LR_CHUNK)).astype(np.float32)
    x = torch.from_numpy(np.ascontiguousarray(x[:sample_chunks]))
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            y = net(x)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    assert tuple(y.shape) == (sample_chunks, 80, 80, 288, 4)
    vox = sample_chunks * int(np.prod(LR_CHUNK[:3]))
    if want_output:
        return vox, times, cores, y.numpy()
    return vox, times, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 12))       # bounded sample: about 10 - 20 s of host work
    warm = 1 if args.warmup > 0 else 0
    vox, times, cores = cpu_reference(steps, warm)
    ms = 1e3 * float(np.mean(times))
    value = vox / (ms / 1e3)
    sample = (f"{steps} timed pass(es) ({sum(times):.1f} s) of 1 LR chunk {LR_CHUNK} through the "
              "torch-CPU port of the literal reference op order (pad3 -> conv -> crop2 per layer, "
              "oneDNN, fp32)")
    line = {"metric": METRIC, "value": value, "unit": "LR voxels/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference", "config": {"workload": WORKLOAD, "chunks_per_step": 1},
            "cpu_baseline": {"value": value, "unit": "LR voxels/s", "cores": cores,
                             "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "LR voxels/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chunks", type=int, default=37,
                    help="LR chunks per device-timed step (37 chunks = 36.0 ring-kernel work items "
                         "per SM on 148 SMs; any smaller batch leaves 2.7 %% of a wave idle)")
    ap.add_argument("--e2e-chunks", type=int, default=8,
                    help="chunks per batch (pass_workers) of the end-to-end ForwardPass.run domain")
    ap.add_argument("--precision", default="fp16c", choices=["fp16c", "bf16", "bf16x3", "fp32"])
    ap.add_argument("--impl", default="sup3r_b200", choices=["sup3r_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-mode", action="store_true")
    ap.add_argument("--no-train-step", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from sup3r_b200 import ops, parallel
    from sup3r_b200.models import Sup3rGan
    from sup3r_b200 import configs as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa_cpus = None
    if world > 1:
        # pinned host buffers on the GPU's NUMA node (before any pinned allocation)
        numa_cpus = parallel.bind_to_gpu_numa(local)
        numa_cpus = len(numa_cpus) if numa_cpus else None
        parallel.init_from_env("nccl")
    dev = torch.device("cuda", local)
    W = max(args.warmup, 3)
    K = args.steps
    B = args.chunks

    hl = gen_config()
    Sup3rGan.seed(0)
    model = Sup3rGan(hl, C.discriminator(3, "same", (2048, 1024)), precision=args.precision,
                     default_device=f"/gpu:{local}",
                     meta={"lr_features": ["u_10m", "v_10m", "u_100m", "v_100m"],
                           "hr_out_features": ["u_10m", "v_10m", "u_100m", "v_100m"],
                           "s_enhance": S_ENH, "t_enhance": T_ENH})
    model.generator.build((B, *LR_CHUNK))
    if world > 1:
        parallel.broadcast_weights([model.generator])   # the one collective: weights, once
    rng = np.random.default_rng(42 + rank)
    x_host = torch.from_numpy(rng.standard_normal((B, *LR_CHUNK)).astype(np.float32)).pin_memory()
    x_dev = x_host.to(dev)
    plan = model.plan_for(model.generator, args.precision)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step_device():
        return plan.run_graphed(x_dev)

    for _ in range(W):
        out = step_device()
    torch.cuda.synchronize()
    assert tuple(out.shape) == (B, 80, 80, 288, 4)

    def barrier():
        if world > 1:
            dist.barrier()

    # ---------------- device-timed value -------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    torch.cuda.synchronize()
    l0 = ops.launch_count()
    evs = []
    for _ in range(K):
        flush.zero_()                       # L2 flush between timed iterations (outside events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_device()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    barrier()
    launches = ops.launch_count() - l0
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    clocks = sampler.finish() if sampler else None
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    vox_step = B * int(np.prod(LR_CHUNK[:3]))
    value = world * vox_step * K / (total_ms / 1e3)

    # ---------------- end to end: host buffers in, host buffers out ------------------------------
    # The reference-facing call: ForwardPass.run(strategy, node_index=rank) on a synthetic LR
    # domain of world x E2E_BATCHES x B chunks, sharded by strategy.node_chunks (the reference's
    # np.array_split over nodes, sup3r/pipeline/strategy.py:363-372; no data-path collective).
    # Timed region per rank: lazy host chunking, pinned H2D, generator, device-side output check,
    # D2H of every chunk's result into its own pinned host buffer.  Results leave the GPU as
    # float16 (what the reference's writers store are scaled 16-bit integers, sup3r/utilities/
    # output_attrs.json): the fp32 figure is reported next to it at every N.
    from sup3r_b200.pipeline import ArrayInputHandler, ForwardPass, ForwardPassStrategy
    from sup3r_b200.pipeline.engine import GeneratePipeline
    n_e2e = min(K, E2E_BATCHES)
    Be = args.e2e_chunks
    dom = (LR_CHUNK[0] * Be, LR_CHUNK[1] * world, LR_CHUNK[2] * n_e2e)
    dom_data = np.random.default_rng(7).standard_normal((*dom, 4)).astype(np.float32)
    feats = model.lr_features

    def fwp_run(out_dtype):
        strat = ForwardPassStrategy(model=model, input_handler=ArrayInputHandler(dom_data, feats),
                                    fwp_chunk_shape=LR_CHUNK[:3], spatial_pad=0, temporal_pad=0,
                                    pass_workers=Be, max_nodes=world, output_dtype=out_dtype)
        assert len(strat.node_chunks) == world and len(strat.node_chunks[rank]) == Be * n_e2e
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = ForwardPass.run(strat, rank)
        torch.cuda.synchronize()
        dt_s = time.perf_counter() - t0
        assert len(outs) == Be * n_e2e and next(iter(outs.values())).shape == (80, 80, 288, 4)
        nbytes = sum(o.nbytes for o in outs.values())
        del outs
        tt = torch.tensor([dt_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(np.prod(dom)) / float(tt.item()), nbytes // n_e2e

    e2e_runs = {}
    for od in ("float16", "float32"):
        fwp_run(od)                                  # builds the pipeline (graphs, pinned slots)
        best = max(fwp_run(od) for _ in range(2))
        e2e_runs[od] = best
    e2e_value, d2h = e2e_runs["float16"]
    h2d = Be * int(np.prod(LR_CHUNK)) * 4
    del dom_data
    # the same chunks through the bare pinned pipeline (no tiler), fp32 results, for comparison
    pipe = GeneratePipeline(model, (Be, *LR_CHUNK), precision=args.precision)
    x_np = rng.standard_normal((Be, *LR_CHUNK)).astype(np.float32)
    vox_e2e = Be * int(np.prod(LR_CHUNK[:3]))
    for y_np in pipe.run([x_np] * 3):
        pass
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for y_np in pipe.run([x_np] * K):
        pass
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    pipe_value = world * vox_e2e * K / float(te.item())
    # the plain synchronous call a user makes (numpy in -> numpy out), for comparison
    for _ in range(2):      # warm the pinned-host allocator cache (two result buffers alternate)
        y_sync = model.generate(x_np, precision=args.precision)
    t0 = time.perf_counter()
    for _ in range(5):
        y_sync = model.generate(x_np, precision=args.precision)
    sync_value = vox_e2e * 5 / (time.perf_counter() - t0)
    del y_sync

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (tcgen05 body conv) ----------------------
    peaks, peaks_src = load_peaks()
    flops_chunk = algorithmic_flops_per_chunk(hl, LR_CHUNK)
    roof = None
    try:
        # the body convolution as the model runs it: 64 -> 64 channels, 3x3x3, 16-bit padded in /
        # out; 17 launches per step write LeakyReLU output only ("plain"), 16 add the
        # SkipConnection pair and write hi + lo ("residual")
        n, dims = B, (16, 16, 288)
        xb = torch.randn((n, *dims, 64), device=dev)
        wb = torch.randn((3, 3, 3, 64, 64), device=dev) * 0.03
        bb = torch.randn(64, device=dev) * 0.1
        split = args.precision in ("bf16x3", "fp16c")
        c_mode = args.precision == "fp16c"
        fmt = 2 if c_mode else 0
        x_hi, x_lo = ops.pack_act_pad16(xb, split=split, fmt=fmt)
        packed = ops.pack_weights_umma(wb, split=split, ndim=3, fmt=fmt)
        w_hi, w_lo = packed[:2]
        acc_scale = packed[2] if len(packed) == 3 else 0.0
        r_hi, r_lo = ops.pack_act_pad16(torch.randn_like(xb), split=True, fmt=fmt)
        y_hi = torch.empty_like(x_hi)
        y_lo = torch.empty_like(x_hi)

        def time_variant(residual):
            spec = ops.ConvSpec(3, 64, 64, (3, 3, 3), pad_lo=(1, 1, 1), pad_hi=(1, 1, 1),
                                pad_mode=1, act=0 if residual else 2, alpha=0.2)

            def body():
                if residual and (c_mode or not split):
                    ops.conv_fwd_umma(x_hi, x_lo, w_hi, w_lo, bb, spec, n, dims, want_f32=False,
                                      out_hi=y_hi, out_lo=y_lo, res_hi=r_hi, res_lo=r_lo,
                                      fmt=fmt, acc_scale=acc_scale)
                else:
                    ops.conv_fwd_umma(x_hi, x_lo, w_hi, w_lo, bb, spec, n, dims, want_f32=False,
                                      out_hi=y_hi, out_lo=y_lo if split else None, fmt=fmt,
                                      acc_scale=acc_scale)
            for _ in range(3):
                body()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = (torch.cuda.Event(enable_timing=True),
                          torch.cuda.Event(enable_timing=True))
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return float(np.mean(ts))
        ms_plain, ms_res = time_variant(False), time_variant(True)
        k_flops = 2.0 * n * np.prod(dims) * 27 * 64 * 64
        n_plain, n_res = 17, 16
        k_ms = (n_plain * ms_plain + n_res * ms_res) / (n_plain + n_res)
        achieved = k_flops / (k_ms / 1e3) / 1e12
        # the kernel is timed ALONE (synchronised graph replays): the burst figure is its peak
        peak = peaks["bf16_tflops"]
        peak_sus = peaks.get("bf16_tflops_sustained") or peak
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r02_body_conv_traffic.json" if c_mode
                          else "r01_body_conv_traffic.json")
        traffic_note = None
        if os.path.exists(tp):
            # ncu --set full capture of a launch over 8 chunks (profiles/): DRAM bytes are
            # linear in the voxels of a launch -> scaled to this launch's n chunks
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            if traffic is not None:
                traffic = float(traffic) * n / 8.0
                traffic_note = ("dram__bytes_read + write of the 8-chunk ncu capture x n / 8 "
                                "(launch mix 17 plain + 16 residual)")
        kname = ("conv_umma_zring_kernel<4, EPI_V4> (fp16 pass + e4m3 correction pass)" if c_mode
                 else ("conv_umma_tile_kernel (split operands, 3 MMA passes)" if split
                       else "conv_umma_zring_kernel<4, EPI_V4>"))
        roof = {"bound": "tensor", "kernel": kname + " (64->64 3x3x3 reflect "
                f"conv, {n}x16x16x288 voxels, 16-bit padded in/out; launch mix of the model step: "
                f"{n_plain} plain + {n_res} residual launches)", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": traffic_note,
                "algorithmic_bytes_per_launch": float(n) * np.prod(dims) * (
                    (256 + 256) * n_plain + (256 + 256 + 256) * n_res) / (n_plain + n_res)
                if c_mode else None,
                "peak_source": f"{peaks_src} bf16_tflops (burst: kernel timed in isolation)",
                "frac_of_sustained_peak": achieved / peak_sus, "kernel_ms": k_ms,
                "kernel_ms_plain": ms_plain, "kernel_ms_residual": ms_res,
                "achieved_plain": k_flops / (ms_plain / 1e3) / 1e12,
                "achieved_residual": k_flops / (ms_res / 1e3) / 1e12,
                "algorithmic_flops_per_launch": k_flops}
    except Exception as e:  # pragma: no cover
        roof = {"error": repr(e)[:300]}

    # ---------------- the other precision modes, device-timed on the same input ---------------
    # (scheme table: which operand scheme tops out where; errors against the CPU oracle below)
    modes = None
    outs_chunk0 = {args.precision: plan.run(x_dev[:1]).cpu().numpy()}
    if not args.no_parity_mode:
        modes = {}
        for pm in ("bf16", "bf16x3", "fp16c"):
            if pm == args.precision:
                modes[pm] = {"value": value / world, "ms_per_step": total_ms / K}
                continue
            try:
                plan_m = model.plan_for(model.generator, pm)
                for _ in range(2):
                    plan_m.run_graphed(x_dev)
                torch.cuda.synchronize()
                ts = []
                for _ in range(5):
                    flush.zero_()
                    e0, e1 = (torch.cuda.Event(enable_timing=True),
                              torch.cuda.Event(enable_timing=True))
                    e0.record()
                    plan_m.run_graphed(x_dev)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                modes[pm] = {"value": vox_step / (float(np.mean(ts)) / 1e3),
                             "ms_per_step": float(np.mean(ts))}
                outs_chunk0[pm] = plan_m.run(x_dev[:1]).cpu().numpy()
            except Exception as e:  # pragma: no cover
                modes[pm] = {"error": repr(e)[:200]}

    cpu, parity = None, None
    if not args.no_cpu_baseline and world == 1:
        # the CPU port runs the GPU arm's weights on its first chunk: its timing is the
        # cpu_baseline, its fp32 output the parity oracle (checker only, never the product path)
        n_cpu = 8                         # about 10 s of host work on the box's 16 cores
        vox, times, cores, y_ref = cpu_reference(n_cpu, 1, weights=model.generator.get_weights(),
                                                 x=x_host.numpy(), want_output=True)
        cpu = {"value": vox / float(np.mean(times)), "unit": "LR voxels/s", "cores": cores,
               "kind": "port",
               "sample": f"{n_cpu} timed passes (after 1 warm-up, {sum(times):.1f} s) of 1 LR "
                         "chunk 16x16x24x4 through the torch-CPU port of the literal reference "
                         "op order (fp32, oneDNN)"}
        y_ref = y_ref.astype(np.float64)
        sc, rms = np.abs(y_ref).max(), np.sqrt(np.mean(y_ref ** 2))
        parity = {"oracle": "oracle/torch_ref.py (fp32 CPU port of the literal reference layer "
                            "sequence) on the same weights and LR chunk 0", "tolerance": 1e-3}
        for pm, y in outs_chunk0.items():
            d = np.abs(y.astype(np.float64) - y_ref)
            parity[pm] = {"max_rel": float(d.max() / sc),
                          "rms_rel": float(np.sqrt(np.mean(d ** 2)) / rms),
                          "median_elementwise_rel": float(np.median(d / np.maximum(np.abs(y_ref), 1e-30)))}
        # the e2e path delivers the result as float16: its rounding on top of the kernels' error
        y16 = outs_chunk0[args.precision].astype(np.float16).astype(np.float64)
        d16 = np.abs(y16 - y_ref)
        parity["e2e_float16_output"] = {"max_rel": float(d16.max() / sc),
                                        "rms_rel": float(np.sqrt(np.mean(d16 ** 2)) / rms)}
        parity["headline_within_tolerance"] = bool(
            parity["e2e_float16_output"]["max_rel"] < 1e-3 and
            parity[args.precision]["max_rel"] < 1e-3 and parity[args.precision]["rms_rel"] < 1e-3)

    train = None
    if not args.no_train_step:
        try:
            del pipe
            torch.cuda.empty_cache()
            train = train_step_block(dev, peaks)
        except Exception as e:  # pragma: no cover
            train = {"error": repr(e)[:300]}

    line = {
        "metric": METRIC, "value": value, "unit": "LR voxels/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"bf16": "bf16", "bf16x3": "bf16x3", "fp32": "f32",
                                       "fp16c": "f16+e4m3 (fp32 accumulate)"}[
            args.precision], "data": "synthetic",
        "config": {"workload": WORKLOAD, "chunks_per_step": B, "e2e_chunks_per_batch": Be,
                   "precision": args.precision,
                   "l2": "flushed between timed steps (256 MiB write outside the event pairs)",
                   "parallelism": f"chunk-dp{world}", "cuda_graph": True},
        "algorithmic_tflops": flops_chunk * B * K * world / (total_ms / 1e3) / 1e12,
        "frac_of_bf16_peak": (flops_chunk * B * K * world / (total_ms / 1e3) / 1e12)
        / (world * (peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"])),
        "e2e": {"value": e2e_value, "unit": "LR voxels/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": int(d2h),
                "api": "ForwardPass.run(strategy, node_index=rank): LR domain "
                       f"{list(dom)} = {world} x {n_e2e} batches of {Be} chunks sharded by "
                       "strategy.node_chunks; host chunking, pinned H2D / D2H, device-side "
                       "output check; float16 results in host memory; wall clock, max over "
                       "ranks, best of 2 after 1 warm-up run",
                "out_dtype": "float16", "batches_per_rank": n_e2e,
                "frac_of_device_rate": e2e_value / value,
                "fp32_out_value": e2e_runs["float32"][0],
                "fp32_out_d2h_bytes_per_step": int(e2e_runs["float32"][1]),
                "generate_pipeline_fp32_value": pipe_value,
                "sync_generate_value": sync_value, "numa_cpus": numa_cpus},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "parity": parity, "precision_modes": modes, "train_step": train,
        "host_cores": os.cpu_count(),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
