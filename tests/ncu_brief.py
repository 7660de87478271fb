"""Diagnostics: the handful of ncu metrics worth reading first, from a .ncu-rep (last kernel of the report)."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps', 'launch__occupancy_limit_blocks',
        'launch__grid_size', 'launch__block_size', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_pipe_lsu.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum']
for path in sys.argv[1:]:
  out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units, vals = rows[0], rows[1], rows[-1]
  print("==", path, vals[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else "")
  for w in WANT:
    if w in hdr:
      print("  %-62s %s %s" % (w, vals[hdr.index(w)], units[hdr.index(w)]))
  st = [(float(vals[i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and vals[i]]
  for v, h in sorted(st, reverse=True)[:7]:
    print("  stall %-40s %.2f" % (h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], v))
