#!/usr/bin/env python
"""Summarise .ncu-rep files (read here, no GPU needed) into a markdown table: per profiled launch the duration, DRAM
bytes / throughput, tensor-pipe activity, issue utilisation, occupancy, registers.
    python tools/ncu_summary.py gpurun_out/r2_gemm.ncu-rep [...] > profiles/r2/ncu_full.md"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active cycles)"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor instr"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem")]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"## {path}: no data\n")
            continue
        hdr, units = rows[0], rows[1]
        print(f"## `{path}`\n")
        print("| launch | kernel | " + " | ".join(n for _, n in KEYS) + " |")
        print("|---|---|" + "---:|" * len(KEYS))
        for i, r in enumerate(rows[2:]):
            name = r[hdr.index("Kernel Name")].split("(")[0][-48:]
            cells = []
            for k, _ in KEYS:
                if k in hdr:
                    j = hdr.index(k)
                    v = r[j]
                    try:
                        v = f"{float(v.replace(',', '')):.4g}"
                    except ValueError:
                        pass
                    cells.append(f"{v} {units[j]}".strip())
                else:
                    cells.append("-")
            print(f"| {i} | `{name}` | " + " | ".join(cells) + " |")
        print()


if __name__ == "__main__":
    main()
