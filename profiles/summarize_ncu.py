#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + SASS sample distribution) into a small text file for profiles/."""
import collections
import csv
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__occupancy_limit', 'launch__waves', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'Kernel Name', 'launch__shared_mem_per_block_dynamic',
        'lts__t_sector_hit_rate', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red', 'lts__throughput.avg.pct',
        'smsp__inst_executed_pipe_fp64', 'sm__inst_executed_pipe_lsu']


def main(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = [f"# {title}", f"# source: {rep} (scratch); selected raw metrics of the first captured launch"]
    for h, u, v in zip(hdr, units, vals):
        if any(k in h for k in KEEP):
            lines.append(f"{h} [{u}] = {v}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    data = []
    for r in rows[2:]:  # several launches are concatenated: keep the first table
        if len(r) != len(hdr):
            break
        data.append(r)
    iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    tot = sum(int(r[iS] or 0) for r in data)
    lines.append(f"# SASS sample distribution ({tot} samples, {len(data)} instructions), blocks of 100 instructions")
    for b in range(0, len(data), 100):
        blk = data[b:b + 100]
        smp = sum(int(r[iS] or 0) for r in blk)
        ex = sum(int(r[iEx] or 0) for r in blk)
        ops = collections.Counter((r[iSrc].split()[1] if r[iSrc].startswith('@') else r[iSrc].split()[0]) for r in blk)
        lines.append(f"sass[{b:4d}:{b+100:4d}] samples {100*smp/max(tot,1):5.1f}%  executed {ex:>11d}  top ops {ops.most_common(5)}")
    top = sorted(range(len(data)), key=lambda k: -int(data[k][iS] or 0))[:12]
    lines.append("# hottest instructions (index, samples, executed, SASS)")
    for k in sorted(top):
        lines.append(f"  {k:5d} {data[k][iS]:>7} {data[k][iEx]:>11} {data[k][iSrc][:80]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
