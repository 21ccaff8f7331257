"""Summarise an `ncu --page raw --csv` dump: key counters and top stall reasons per kernel."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum']
idx = [hdr.index(w) for w in want if w in hdr]
st = [i for i, h in enumerate(hdr) if 'smsp__average_warp' in h and 'per_issue_active' in h and 'not_issued' not in h]
for r in rows[2:]:
    print('---')
    for i in idx:
        print(f"  {hdr[i]:70s} {r[i]:>20s} {units[i]}")
    top = sorted([(float(r[i]), hdr[i].split('issue_stalled_')[-1].split('_per')[0]) for i in st if r[i]], reverse=True)[:6]
    print('  stalls (warp cycles per issue):', ', '.join(f"{n}={v:.2f}" for v, n in top))
