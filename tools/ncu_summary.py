#!/usr/bin/env python
"""Summarise ncu output into the small text files kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/x/launches.csv  > profiles/r01_launches.md
    python tools/ncu_summary.py full     gpurun_out/x/prof.ncu-rep  > profiles/r01_gemm_full.md

`launches`: the CSV written by  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...
`full`    : a .ncu-rep captured with --set full; read through `ncu -i ... --page raw --csv`.
"""
import collections
import csv
import subprocess
import sys

KEY_METRICS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.max",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tc.sum",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static",
    "launch__grid_size",
    "launch__block_size",
    "launch__cluster_size",
    "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
]


def _rows(path):
    return [r for r in csv.reader(open(path, newline="")) if len(r) > 5]


def launches(path, last_n=None):
    rows = _rows(path)
    hdr, rows = rows[0], rows[1:]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0][:100]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    print(f"# ncu launch list: {path}\n")
    print(f"{len(rows)} launches, {total / 1e3:.1f} us of kernel time in total (cold-cache, serialised; shares are what matter)\n")
    print("| share | total us | launches | avg us | kernel |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * t / total:.1f}% | {t / 1e3:.1f} | {n} | {t / n / 1e3:.1f} | `{k}` |")
    if last_n:
        print(f"\n## last {last_n} launches (one step, in issue order)\n")
        print("| us | grid | block | kernel |\n|---|---|---|---|")
        for r in rows[-last_n:]:
            print(f"| {float(r[vi].replace(',', '')) / 1e3:.1f} | {r[gi]} | {r[bi]} | `{r[ki].split('(')[0][:100]}` |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    hdr, units, rows = rows[0], rows[1], rows[2:]
    print(f"# ncu --set full summary: {path}\n")
    ki = hdr.index("Kernel Name")
    for n, r in enumerate(rows):
        print(f"## launch {n}: `{r[ki][:120]}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for m in KEY_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {m} | {r[i]} | {units[i]} |")
        # derived
        try:
            t_us = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
            u = units[hdr.index("gpu__time_duration.sum")]
            scale = {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
            t_us *= scale

            def mb(name):
                i = hdr.index(name)
                v = float(r[i].replace(",", ""))
                return v * {"Mbyte": 1.0, "Gbyte": 1e3, "Kbyte": 1e-3, "byte": 1e-6}.get(units[i], 1.0)
            traffic = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
            print(f"| (derived) dram traffic read+write | {traffic:.1f} | Mbyte |")
            print(f"| (derived) dram GB/s | {traffic / t_us * 1e3:.1f} | GB/s |")
        except Exception:
            pass
        print()


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    if mode == "launches":
        launches(path, int(sys.argv[3]) if len(sys.argv) > 3 else None)
    else:
        full(path)
