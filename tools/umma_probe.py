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
    # name, shape (n,z,y,x), ndim, cout, tuning dict, split, extra spec kwargs
    ("ring_base", (1, 4, 16, 16), 3, 64, dict(), 0, {}),
    ("ring_cols", (2, 16, 16, 40), 3, 64, dict(max_ctas=3), 0, {}),
    ("ring_cols7", (2, 16, 16, 40), 3, 64, dict(max_ctas=7), 0, {}),
    ("ring_z5_b2", (2, 5, 16, 24), 3, 64, dict(max_ctas=2), 0, {}),
    ("ring_z7_y20", (1, 7, 20, 10), 3, 64, dict(max_ctas=2), 0, {}),
    ("ring_r2", (1, 7, 12, 10), 3, 64, dict(tiles=2, max_ctas=1), 0, {}),
    ("ring_r1", (1, 3, 20, 9), 3, 64, dict(tiles=1), 0, {}),
    ("ring_p6", (1, 9, 16, 16), 3, 64, dict(ring_slots=6, max_ctas=1), 0, {}),
    ("ring_ws3", (1, 9, 16, 16), 3, 64, dict(w_stages=3, max_ctas=2), 0, {}),
    ("ring_rep3", (1, 4, 16, 8), 3, 64, dict(), 0, dict(out_repeat=(1, 1, 3))),
    ("ring_c72", (1, 6, 16, 16), 3, 72, dict(max_ctas=1), 0, {}),
    ("ring_c48", (1, 6, 16, 16), 3, 48, dict(max_ctas=1), 0, {}),
    ("zcat_y20", (1, 4, 16, 16), 3, 64, dict(scheme=1), 0, {}),
    ("zcat_y18", (1, 4, 16, 16), 3, 64, dict(box_y=18, scheme=1), 0, {}),
    ("zcat_z5_b2", (2, 5, 16, 24), 3, 64, dict(scheme=1), 0, {}),
    ("zcat_r2", (1, 7, 12, 10), 3, 64, dict(tiles=2, scheme=1), 0, {}),
    ("zcat_r1", (1, 3, 20, 9), 3, 64, dict(tiles=1, scheme=1), 0, {}),
    ("zcat_split", (1, 6, 16, 16), 3, 64, dict(), 1, {}),
    ("zcat_rep3", (1, 4, 16, 8), 3, 64, dict(scheme=1), 0, dict(out_repeat=(1, 1, 3))),
    ("zcat_c72", (1, 4, 16, 16), 3, 72, dict(scheme=1), 0, {}),
    ("tile_2d", (3, 1, 20, 20), 2, 64, dict(), 0, {}),
    ("tile_2d_big", (2, 1, 40, 36), 2, 64, dict(), 1, {}),
    ("tile_head200", (1, 4, 16, 16), 3, 200, dict(), 0, dict(d2s=5)),
    ("tile_c128", (1, 4, 10, 16), 3, 128, dict(), 0, {}),
]


def run_case(name, shape, ndim, cout, tune, split, extra):
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
    spec = ops.ConvSpec(ndim, 64, cout, k, pad_lo=pad, pad_hi=pad, pad_mode=1, act=2, alpha=0.2,
                        **extra)
    # reference on bf16-rounded operands with the exact fp32 kernel
    x_hi, x_lo = ops.pack_act_pad16(xt, split=bool(split))
    w_hi, w_lo = ops.pack_weights_umma(w, split=bool(split), ndim=ndim)
    xr = ops.unpack_act_pad16(x_hi, x_lo, ndim)
    # reference weights = the 16-bit rounded values (hi [+ lo]) in keras layout
    wb = w.to(torch.bfloat16)
    wr = wb.float()
    if split:
        wr = wr + (w - wr).to(torch.bfloat16).float()
    ref = ops.conv_fwd(xr, wr, b, spec)
    t = UmmaTuning(**tune)
    dims = (z, y, x) if ndim == 3 else (1, y, x)
    plain = spec.d2s == 1 and spec.d2t == 1
    res_t = torch.randn_like(ref) if plain and not extra else None
    if res_t is not None:
        ref = ref + res_t
    out, out_hi, out_lo = ops.conv_fwd_umma(x_hi, x_lo, w_hi, w_lo, b, spec, n, dims,
                                            residual=res_t, want_pad16=plain, tune=t)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    res = dict(name=name, max_err=err, ref_scale=scale)
    if plain:
        back = ops.unpack_act_pad16(out_hi, out_lo, ndim)
        res["err_pad16"] = (back - ref).abs().max().item()
        # mirrored halo check: repack the fp32 result and compare the whole padded tensor
        ref_hi, _ = ops.pack_act_pad16(out)
        res["halo_err"] = (ref_hi.float() - out_hi.float()).abs().max().item()
    return res


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "inproc":
        # every case whose name starts with the prefix, in this process (fast; a trap ends it)
        for c in CASES:
            if c[0].startswith(sys.argv[2]):
                try:
                    print(run_case(*c), flush=True)
                except Exception as e:  # noqa
                    print(dict(name=c[0], error=repr(e)[:500]), flush=True)
                    break
        sys.exit(0)
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
