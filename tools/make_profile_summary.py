"""Turn ncu artefacts in gpurun_out/ into small tracked summaries under profiles/.

  python tools/make_profile_summary.py r01
    gpurun_out/<tag>_launches_raw.csv   -> profiles/<tag>_launches.md  (+ per-kernel share)
    gpurun_out/<tag>_*_full.ncu-rep     -> profiles/<tag>_<name>_full.md (key raw metrics)
"""
import collections
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__pipe_tensor_cycles_active",
        "sm__inst_executed_pipe_tensor", "dram__bytes_read.sum [", "dram__bytes_write.sum [",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum [",
        "lts__throughput.avg.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct", "sm__warps_active.avg.pct_of_peak", "launch__registers_per_thread [",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__inst_executed.sum [", "smsp__inst_executed.sum [", "sass__inst_executed_local",
        "sm__mem_tensor_cycles_active.avg.pct", "smsp__pipe_fma_cycles_active.avg.pct",
        "sm__inst_executed_pipe_fma", "l1tex__data_bank_conflicts_pipe_lsu.sum ["]


def launches(tag):
    fp = os.path.join(ROOT, "gpurun_out", f"{tag}_launches_raw.csv")
    if not os.path.exists(fp):
        return
    rows = list(csv.reader(open(fp)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ik, iv, ig, ib = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Grid Size",
                                              "Block Size"))
    agg = collections.OrderedDict()
    for r in data:
        name = r[ik].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, r[ig], r[ib]])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) / 1e3
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if k.startswith("s3::"))
    out = [f"# {tag}: ncu launch list of `python bench.py --steps 2 --warmup 3`",
           "", "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised "
           "launches: compare SHARES, not absolutes).  All launches of the process (warm-up, "
           "graph capture, timed steps, e2e pipeline, roofline loop).", "",
           f"total {len(data)} launches, {tot / 1e3:.2f} ms; kernels of libsup3r_b200 (`s3::`): "
           f"{100 * ours / tot:.1f} % of device time", "",
           "| kernel | launches | total us | share | avg us | grid | block |", "|---|---|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k[:70]}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % | "
                   f"{v[1] / v[0]:.1f} | {v[2]} | {v[3]} |")
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")
    print("wrote", f"profiles/{tag}_launches.md")


def full(tag):
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_*_full.ncu-rep"))):
        name = os.path.basename(rep).replace(".ncu-rep", "")
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        out = [f"# {name}: `ncu --set full --clock-control none --import-source on` ({len(rows) - 2} launch(es); SASS hot spots: first launch)", ""]
        summary = {}
        traffic_rows = []
        for r in rows[2:]:
            out += [f"kernel `{r[4][:100]}` grid {r[8]} block {r[7]}", "",
                    "| metric | unit | value |", "|---|---|---|"]
            for h, u, v in zip(hdr, units, r):
                if any(k in f"{h} [" if k.endswith("[") else k in h for k in KEYS):
                    out.append(f"| {h} | {u} | {v} |")
                    summary[h] = v
            rd = float(summary.get("dram__bytes_read.sum", "0").replace(",", "") or 0)
            wr = float(summary.get("dram__bytes_write.sum", "0").replace(",", "") or 0)
            ui = hdr.index("dram__bytes_read.sum") if "dram__bytes_read.sum" in hdr else None
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(
                units[ui] if ui is not None else "byte", 1.0)
            out += ["", f"dram traffic per launch: {(rd + wr) * scale / 1e6:.2f} MB"]
            traffic_rows.append({"dram_bytes_per_launch": (rd + wr) * scale, "kernel": r[4][:80],
                                 "gpu_time_us": summary.get("gpu__time_duration.sum")})
        json.dump(traffic_rows if len(traffic_rows) > 1 else traffic_rows[0],
                  open(os.path.join(ROOT, "profiles", f"{name}_traffic.json"), "w"), indent=1)
        # hottest SASS lines by stall samples
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source",
                              "sass"], capture_output=True, text=True).stdout
        srows = list(csv.reader(src.splitlines()))
        if len(srows) > 3:
            shdr = srows[1]
            isamp, isrc, iex = (shdr.index(k) for k in ("# Samples", "Source",
                                                         "Instructions Executed"))
            data = [r for r in srows[2:] if len(r) > max(isamp, isrc, iex)]
            # (multi-launch reports repeat the header block per launch: keep the first launch)
            for k, r in enumerate(data):
                if r[isamp] == '# Samples' or not (r[isamp] or '0').isdigit():
                    data = data[:k]
                    break
            tot = sum(int(r[isamp] or 0) for r in data)
            out += ["", f"## hottest SASS instructions ({tot} samples over {len(data)} instrs)", "",
                    "| idx | samples | executed | instruction | top stalls |", "|---|---|---|---|---|"]
            scols = [i for i, h in enumerate(shdr) if h.startswith("stall_") and "Not" not in h]
            for i in sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:25]:
                r = data[i]
                st = {shdr[c][6:]: int(r[c]) for c in scols if r[c] and int(r[c]) > 0}
                st = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
                out.append(f"| {i} | {r[isamp]} | {r[iex]} | `{r[isrc].strip()[:70]}` | {st} |")
            mn = [r[isrc] for r in data]
            flags = {k: sum(k in m for m in mn) for k in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR",
                                                          "HMMA", "FFMA", "LDGSTS")}
            out += ["", f"SASS mnemonic counts: {flags}"]
        open(os.path.join(ROOT, "profiles", f"{name}.md"), "w").write("\n".join(out) + "\n")
        print("wrote", f"profiles/{name}.md")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launches(tag)
    full(tag)
