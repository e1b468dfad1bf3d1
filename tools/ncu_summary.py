#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.csv
  python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep profiles/r01_pileup_raw.csv
"""
import collections
import csv
import io
import subprocess
import sys

RAW_METRICS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
               "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
               "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
               "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
               "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
               "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
               "smsp__inst_executed_op_shared_atom.sum", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
               "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_elapsed.max"]


def launches(src, dst):
    lines = [l for l in open(src).read().splitlines() if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO("\n".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    # libcratac's kernels: "cratac::k_..." in a full launch list, bare "k_..." when the capture was filtered with -k regex:^k_
    ours = {k: v for k, v in agg.items() if "cratac::" in k or k.startswith("k_") or k.startswith("void k_")}
    tot = sum(v[1] for v in ours.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: libcratac kernels only (torch data-generation kernels dropped)\n")
        f.write("kernel,launches,total_us,share_of_cratac_kernels\n")
        for k, (n, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.4f\n" % (k, n, t, t / tot))


def raw(src, dst):
    out = subprocess.check_output(["ncu", "-i", src, "--page", "raw", "--csv"]).decode()
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        w = csv.writer(f)
        cols = [m for m in RAW_METRICS if m in idx]
        w.writerow(cols)
        w.writerow([units[idx[m]] for m in cols])
        for r in rows[2:]:
            w.writerow([r[idx[m]] for m in cols])


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
