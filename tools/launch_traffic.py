"""ncu launch list with DRAM bytes -> markdown table + the per-step traffic JSON bench.py reads (roofline.traffic).

    python tools/launch_traffic.py gpurun_out/x/launches.csv <steps captured> <algorithmic bytes per step> profiles/rNN_launches_config3.md profiles/rNN_config3_traffic.json
CSV: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file ...
"""
import collections
import csv
import json
import re
import sys


def main():
    path, steps, alg, md_path, json_path = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), sys.argv[4], sys.argv[5]
    lines = [l for l in open(path) if l.startswith('"')]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = per.setdefault(row["ID"], {"name": row["Kernel Name"]})
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        if row["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        else:
            d["rd" if "read" in row["Metric Name"] else "wr"] = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    fam = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0])
    for d in per.values():
        name = re.sub(r"\(.*", "", d["name"]).replace("tlt::", "")
        f = fam[name]
        f[0] += d["us"]; f[1] += 1; f[2] += d["rd"]; f[3] += d["wr"]
    tot_us = sum(f[0] for f in fam.values())
    rd, wr = sum(f[2] for f in fam.values()), sum(f[3] for f in fam.values())
    with open(md_path, "w") as o:
        o.write(f"{steps} steps captured ({len(per)} launches): {tot_us:.1f} us of kernel time (cold-cache, serialised under ncu), DRAM read {rd / 1e6:.0f} MB + "
                f"write {wr / 1e6:.0f} MB.\nPer step: {(rd + wr) / steps / 1e6:.0f} MB of DRAM traffic against {alg / 1e6:.0f} MB algorithmic "
                f"({(rd + wr) / steps / alg:.2f}x).\n\n| share | total us | launches | DRAM read MB | DRAM write MB | kernel |\n|---|---|---|---|---|---|\n")
        for name, f in sorted(fam.items(), key=lambda kv: -kv[1][0]):
            o.write(f"| {100 * f[0] / tot_us:.1f}% | {f[0]:.1f} | {f[1]} | {f[2] / 1e6:.1f} | {f[3] / 1e6:.1f} | `{name}` |\n")
    gemm_rd = sum(f[2] for n, f in fam.items() if "gemm" in n)
    json.dump({"dram_bytes_per_step": (rd + wr) / steps, "dram_read_bytes_per_step": rd / steps, "dram_write_bytes_per_step": wr / steps,
               "algorithmic_bytes_per_step": alg, "gemm_family_dram_read_bytes_per_step": gemm_rd / steps,
               "source": f"{md_path}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the {len(per) // steps} launches of one step "
                         "(whole step, the scope of roofline.achieved)"}, open(json_path, "w"), indent=1)


if __name__ == "__main__":
    main()
