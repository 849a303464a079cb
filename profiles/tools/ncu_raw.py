#!/usr/bin/env python
"""Key metrics of `ncu -i X.ncu-rep --page raw --csv` (one column per captured launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = """gpu__time_duration.sum launch__grid_size launch__registers_per_thread launch__occupancy_limit_registers
sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed dram__bytes_read.sum dram__bytes_write.sum dram__bytes_read.sum.per_second
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed lts__t_sector_hit_rate.pct lts__t_sectors_srcunit_tex_op_read.sum
lts__t_sectors_srcunit_tex_lookup_hit.sum lts__t_sectors_srcunit_tex_lookup_miss.sum lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
l1tex__throughput.avg.pct_of_peak_sustained_elapsed sm__cycles_elapsed.max""".split()
for w in WANT:
    for i, h in enumerate(hdr):
        if h == w:
            print(f"{w:68s} {units[i]:12s} {[r[i] for r in data]}")
