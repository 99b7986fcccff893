"""Scratch: key metrics of an ncu report as 'name = value unit' lines. usage: ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys, re
txt = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = re.compile(r'^(Kernel Name|dram__bytes_(read|write)\.sum$|gpu__dram_throughput\.avg\.pct|gpu__time_duration\.sum|l1tex__t_sector_hit_rate|'
                  r'l1tex__throughput\.avg\.pct|launch__(block_size|grid_size|registers_per_thread$|occupancy_limit|shared_mem_per_block$)|'
                  r'lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct|lts__t_bytes\.sum$|sass__inst_executed_local|sm__cycles_elapsed\.max|'
                  r'sm__throughput\.avg\.pct|sm__warps_active\.avg\.pct|smsp__average_warps_issue_stalled.*per_issue_active|'
                  r'lts__t_sectors\.sum$|lts__t_sectors_srcunit_tex\.sum$|l1tex__t_sectors_pipe_lsu_mem_global_op_ld\.sum$|smsp__inst_executed\.sum$|smsp__issue_active\.avg\.pct|smsp__thread_inst_executed_per_inst_executed\.ratio|'
                  r'sm__inst_executed_pipe_(alu|fma|lsu|xu|fp64|uniform)\.sum$|l1tex__t_bytes\.sum$|smsp__warps_eligible\.avg\.per_cycle|achieved_occupancy|sm__warps_active)')
for d in data:
    for h, u, v in zip(hdr, units, d):
        if want.match(h): print('%s = %s %s' % (h, v, u))
