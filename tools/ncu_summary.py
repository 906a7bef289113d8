#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key pipe / memory / stall metrics as CSV,
plus the hottest SASS lines.  Usage: ncu_summary.py report.ncu-rep out_prefix"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|sm__cycles_active\.avg|smsp__inst_executed\.sum|"
    r"sm__inst_executed_pipe_(fp64|fma|fmaheavy|alu|xu|lsu|cbu|uniform)\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_(fp64|fma|alu|xu)_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__issue_active\.avg\.pct_of_peak_sustained_elapsed|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active\.ratio|"
    r"l1tex__data_pipe_lsu_wavefronts(_mem_shared)?\.sum(\.pct_of_peak_sustained_elapsed)?|"
    r"l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
    r"smsp__inst_executed_op_shared_atom(_dot_alu|_dot_cas)?\.sum|smsp__inst_executed_op_shared_(ld|st|atom)\.sum|"
    r"l1tex__t_set_accesses_pipe_lsu_mem_shared_op_atom\.sum|"
    r"dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|lts__t_sectors_op_(atom|red)\.sum|"
    r"launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|waves_per_multiprocessor|"
    r"occupancy_limit_[a-z_]+)|sm__maximum_warps_per_active_cycle_pct|smsp__cycles_active\.avg)$")


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    with open(prefix + "_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for name in ("Kernel Name", "Block Size", "Grid Size"):
            if name in hdr:
                w.writerow([name, "", vals[hdr.index(name)]])
        for h, u, v in zip(hdr, units, vals):
            if KEEP.match(h):
                w.writerow([h, u, v])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    i_src, i_smp, i_ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    data = [(int(r[i_smp]), int(r[i_ex]), r[i_src].strip()) for r in rows[2:]
            if len(r) > i_smp and r[i_smp].isdigit()]
    tot = sum(d[0] for d in data) or 1
    with open(prefix + "_hot_sass.txt", "w") as f:
        f.write(f"# {rows[0][1] if len(rows[0]) > 1 else ''}\n# samples {tot}, "
                f"warp instructions {sum(d[1] for d in data)}\n# samples  share  executed  SASS\n")
        for s, e, t in sorted(data, reverse=True)[:40]:
            f.write(f"{s:8d} {100 * s / tot:6.2f}% {e:12d}  {t}\n")
        ops = {}
        for s, e, t in data:
            op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] if t else "?"
            ops[op] = ops.get(op, 0) + e
        f.write("\n# executed warp instructions by opcode\n")
        for op, e in sorted(ops.items(), key=lambda kv: -kv[1])[:30]:
            f.write(f"{e:14d}  {op}\n")


if __name__ == "__main__":
    main()
