#!/usr/bin/env python
"""Per-launch summary of an `ncu --set full` capture exported with `ncu -i x.ncu-rep --page raw --csv`: duration,
tensor-pipe utilisation (sm__pipe_tensor_cycles_active), DRAM bytes, L2 / DRAM / SM throughput, grid, registers."""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor% (active)"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor% (elapsed)"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("launch__grid_size", "grid"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"source: {path}; one line per captured launch")
    print("kernel".ljust(34) + "".join(n.rjust(18) for _, n in COLS if _ in idx))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].replace("void ", "")[:32]
        out = name.ljust(34)
        for key, _ in COLS:
            if key in idx:
                out += f"{r[idx[key]]} {units[idx[key]]}"[:17].rjust(18)
        print(out)


if __name__ == "__main__":
    main(sys.argv[1])
