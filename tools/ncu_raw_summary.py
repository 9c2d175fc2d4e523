#!/usr/bin/env python
"""Text summary of an .ncu-rep (`ncu --set full`) for profiles/: one block per captured launch
with the counters DESIGN.md quotes.  usage: ncu_raw_summary.py <report.ncu-rep> [header line]"""
import csv
import io
import subprocess
import sys

METRICS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__throughput.avg.pct_of_peak_sustained_elapsed l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct
launch__grid_size launch__block_size launch__registers_per_thread launch__shared_mem_per_block_dynamic
launch__shared_mem_per_block_static sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
lts__t_requests_srcunit_tex_op_atom_dot_alu.sum lts__t_requests_srcunit_tex_op_atom_dot_cas.sum
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio""".split()


def main(rep, header=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units, data = rows[0], rows[1], rows[2:]
    ix = {n: i for i, n in enumerate(h)}
    if header:
        print("# " + header)
    print(f"# source report: {rep} (scratch, not committed)")
    for r in data:
        print(f"\n## {r[ix['Kernel Name']][:110]}")
        for m in METRICS:
            if m in ix:
                print(f"{m:95s} {units[ix[m]]:18s} {r[ix[m]]}")


if __name__ == "__main__":
    main(*sys.argv[1:])
