"""Print the key metrics of every kernel in an .ncu-rep (run where ncu is installed)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'smsp__inst_executed_op_global_ld.sum']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') or h.startswith('smsp__average_warp_latency_issue_stalled')]
for r in rows[2:]:
    print('=====', r[idx['Kernel Name']][:100])
    for w in want:
        if w in idx:
            print(f"  {w:75s} {r[idx[w]]} {units[idx[w]]}")
    st = sorted(((float(r[idx[h]].replace(',', '') or 0), h) for h in stall if r[idx[h]]), reverse=True)[:8]
    for v, h in st:
        print(f"  STALL {h:90s} {v:.3f}")
