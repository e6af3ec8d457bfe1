#!/usr/bin/env python3
"""Turn the ncu artefacts in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag> [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof.ncu-rep ...]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second", "smsp__cycles_active.avg",
    "launch__shared_mem_per_block_dynamic", "launch__cluster_size", "sm__inst_executed_pipe_fp64.sum",
    "smsp__inst_executed_pipe_xu.sum", "lts__t_bytes.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
]


def ms(value, unit):
    t = float(value.replace(",", ""))
    return {"ns": t / 1e6, "us": t / 1e3, "usecond": t / 1e3, "ms": t, "msecond": t, "s": t * 1e3, "second": t * 1e3}.get(unit, t)


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms(r["Metric Value"], r["Metric Unit"])
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# per-kernel device time of ONE bench step under `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:hh_`\n")
        f.write("# (serialised, cold caches: compare SHARES, not absolutes)\n")
        f.write("kernel,launches,total_ms,share_pct\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("{},{},{:.3f},{:.1f}\n".format(k, v[0], v[1], 100 * v[1] / tot))
        f.write("TOTAL,{},{:.3f},100.0\n".format(sum(v[0] for v in agg.values()), tot))
    print("wrote", out)


def report(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    with open(out, "w") as f:
        f.write("# `ncu --set full --clock-control none --import-source on` summary of {}\n".format(os.path.basename(rep)))
        for r in rows[2:]:
            f.write("\n## {}\n".format(r[idx["Kernel Name"]]))
            for m in METRICS:
                if m in idx and r[idx[m]] not in ("", "nan", "-nan"):
                    f.write("{:78s} {} {}\n".format(m, r[idx[m]], units[idx[m]]))
            st = []
            for h in stall:
                try:
                    if float(r[idx[h]]) == float(r[idx[h]]):
                        st.append((float(r[idx[h]]), h))
                except ValueError:
                    pass
            for v, h in sorted(st, reverse=True)[:6]:
                f.write("stall {:40s} {:.2f} warps per issue\n".format(
                    h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    print("wrote", out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep", nargs="*", default=[])
    ap.add_argument("--outdir", default="profiles", help="where the summaries go (gpurun_out on the GPU box: only that directory travels back)")
    a = ap.parse_args()
    os.makedirs(a.outdir, exist_ok=True)
    if a.launches:
        launches(a.launches, "{}/{}_launches.csv".format(a.outdir, a.tag))
    for rep in a.rep:
        base = os.path.splitext(os.path.basename(rep))[0]
        base = base[len(a.tag) + 1:] if base.startswith(a.tag + "_") else base
        report(rep, "{}/{}_{}.txt".format(a.outdir, a.tag, base))


if __name__ == "__main__":
    main()
