#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU) into the few metrics the roofline discussion needs.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp_inst"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%")]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of `{path}` (per launch; cold-cache, serialised replays)\n")
    print("| kernel | " + " | ".join(n for _, n in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        vals = []
        for k, _ in KEYS:
            if k in idx:
                v = r[idx[k]].replace(",", "")
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                vals.append(f"{v} {units[idx[k]]}".strip())
            else:
                vals.append("n/a")
        print(f"| {name} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
