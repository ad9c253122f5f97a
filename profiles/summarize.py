"""Summarise gpurun_out/ ncu artefacts into the tracked text files under profiles/.
    python profiles/summarize.py launches gpurun_out/launches_r01.csv > profiles/r01_launches.txt
    python profiles/summarize.py raw gpurun_out/prof_conv_r01.ncu-rep > profiles/r01_conv_ncu.txt
"""
import collections
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "lts__t_bytes.sum",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
        "launch__shared_mem_per_block_dynamic"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"# {path}: {len(data)} launches, {tot / 1e3:.3f} ms total (gpu__time_duration.sum, cold-cache, serialised)")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:12.1f} us {100 * t / tot:6.2f}%  n={n:4d}  avg={t / n:9.1f} us  {k}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}")
    for r in rows[2:]:
        print("---")
        for w in WANT:
            if w in idx:
                print(f"{w}: {r[idx[w]]} {units[idx[w]]}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
