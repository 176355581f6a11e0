"""Hardware probe for the tcgen05 convolution: runs s3_conv_fwd_umma under several descriptor /
tiling variants against the fp32 direct kernel and writes max errors to gpurun_out/umma_probe.json.
Each variant runs in a subprocess so a trapped kernel cannot poison the others."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # name, shape (n,z,y,x), ndim, cout, tuning dict, split
    ("bo0_x10", (1, 4, 16, 16), 3, 64, dict(base_offset_mode=0, box_x=10), 0),
    ("bo1_x10", (1, 4, 16, 16), 3, 64, dict(base_offset_mode=1, box_x=10), 0),
    ("bo0_x16", (1, 4, 16, 16), 3, 64, dict(base_offset_mode=0, box_x=16), 0),
    ("bo1_x16", (1, 4, 16, 16), 3, 64, dict(base_offset_mode=1, box_x=16), 0),
]


def run_case(name, shape, ndim, cout, tune, split):
    import torch
    from sup3r_b200 import ops
    from sup3r_b200._cabi import UmmaTuning
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    n, z, y, x = shape
    xs = (n, z, y, x, 64) if ndim == 3 else (n, y, x, 64)
    xt = torch.randn(xs, device=dev)
    k = (3, 3, 3) if ndim == 3 else (1, 3, 3)
    w = torch.randn(k + (64, cout), device=dev) * 0.05
    b = torch.randn(cout, device=dev) * 0.1
    pad = (1, 1, 1) if ndim == 3 else (0, 1, 1)
    spec = ops.ConvSpec(ndim, 64, cout, k, pad_lo=pad, pad_hi=pad, pad_mode=1, act=2, alpha=0.2)
    # reference on bf16-rounded operands with the exact fp32 kernel
    x_hi, x_lo = ops.pack_act_pad16(xt, split=bool(split))
    w_hi, w_lo = ops.pack_weights_umma(w, split=bool(split))
    xr = ops.unpack_act_pad16(x_hi, x_lo, ndim)
    wr = (w_hi.float() + (w_lo.float() if w_lo is not None else 0))[:, :cout, :]
    wr = wr.permute(0, 2, 1).reshape(k + (64, cout)).contiguous()
    ref = ops.conv_fwd(xr, wr, b, spec)
    t = UmmaTuning(**tune)
    dims = (z, y, x) if ndim == 3 else (1, y, x)
    out, out_hi, _ = ops.conv_fwd_umma(x_hi, x_lo, w_hi, w_lo, b, spec, n, dims, want_pad16=True,
                                       tune=t)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    back = ops.unpack_act_pad16(out_hi, None, ndim)
    err16 = (back - ref).abs().max().item()
    # mirrored halo check: repack the fp32 result and compare the whole padded tensor
    ref_hi, _ = ops.pack_act_pad16(out)
    halo = (ref_hi.float() - out_hi.float()).abs().max().item()
    return dict(name=name, max_err=err, ref_scale=scale, err_pad16=err16, halo_err=halo)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        i = int(sys.argv[1])
        try:
            res = run_case(*CASES[i])
        except Exception as e:  # noqa
            res = dict(name=CASES[i][0], error=repr(e)[:500])
        print("PROBE_RESULT " + json.dumps(res))
        sys.exit(0)
    out = []
    for i, c in enumerate(CASES):
        try:
            r = subprocess.run([sys.executable, __file__, str(i)], capture_output=True, text=True,
                               timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("PROBE_RESULT ")]
            res = json.loads(line[-1][len("PROBE_RESULT "):]) if line else dict(
                name=c[0], error="no result", rc=r.returncode, stderr=r.stderr[-800:])
        except subprocess.TimeoutExpired:
            res = dict(name=c[0], error="timeout")
        print(res, flush=True)
        out.append(res)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "umma_probe.json"), "w"), indent=1)
