#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` exports into the few numbers the roofline discussion uses.
    python tools/ncu_summary.py gpurun_out/prof_conv_raw.csv [...] > profiles/rNN_ncu_full_summary.txt
Also writes a JSON (consumed by bench.py for roofline.traffic) when --json PATH is given."""
import csv
import json
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (hmma) % active"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor memory pipe % elapsed"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("smsp__cycles_elapsed.avg.per_second", "SM clock"),
]


def to_bytes(unit, val):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(val) * m.get(unit, 1)


def main():
    args = sys.argv[1:]
    jpath = None
    if "--json" in args:
        i = args.index("--json")
        jpath = args[i + 1]
        del args[i:i + 2]
    out = {}
    for path in args:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            if len(vals) != len(hdr):
                continue
            d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
            name = d["Kernel Name"][1]
            print(f"== {path}\n   kernel {name}  grid {d['Grid Size'][1]} block {d['Block Size'][1]}")
            for k, label in KEYS:
                if k in d:
                    print(f"   {label:34s} {d[k][1]} {d[k][0]}")
            rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
            us = float(d["gpu__time_duration.sum"][1]) * {"us": 1, "ms": 1e3, "ns": 1e-3}[d["gpu__time_duration.sum"][0]]
            print(f"   dram traffic per launch            {(rd + wr) / 1e6:.2f} MB  ({(rd + wr) / us / 1e3:.0f} GB/s under ncu)")
            out[name.split("(")[0]] = dict(file=path, duration_us=us, dram_read_bytes=rd, dram_write_bytes=wr)
    if jpath:
        json.dump(out, open(jpath, "w"), indent=1)


if __name__ == "__main__":
    main()
