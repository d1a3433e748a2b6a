"""Print the metrics that matter for the raster kernels from an .ncu-rep (first profiled launch)."""
import csv, io, subprocess, sys
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active']
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, u, r = rows[0], rows[1], rows[2]
    print("==", rep)
    for k in KEYS:
        if k in h:
            i = h.index(k); print(f"  {k}: {r[i]} {u[i]}")
    st = sorted(((float(r[i]), k) for i, k in enumerate(h) if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k), reverse=True)
    print("  stalls per issue:", ", ".join(f"{k.split('issue_stalled_')[1].split('_per_')[0]} {v:.2f}" for v, k in st[:8]))
