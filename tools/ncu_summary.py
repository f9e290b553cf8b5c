"""Summarises an .ncu-rep (raw + source pages) into text: key metrics, stall mix, hottest SASS lines."""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']
for vals in rows[2:]:
    print("== kernel:", vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
    for i, h in enumerate(hdr):
        if h in want:
            print("  %-70s %-12s %s" % (h, units[i], vals[i]))
    stalls = [(float(vals[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
              for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
    print("  stalls per issue:", ", ".join("%s=%.2f" % (n, v) for v, n in sorted(stalls, reverse=True) if v > 0.01))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
if hi:
    hdr = rows[hi[0]]
    data = [r for r in rows[hi[0] + 1:] if len(r) == len(hdr)]
    isrc, iex, ismp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    tot = sum(int(r[iex]) for r in data)
    tots = sum(int(r[ismp]) for r in data)
    print("total warp instructions %d, samples %d" % (tot, tots))
    ranked = sorted(range(len(data)), key=lambda k: -int(data[k][ismp]))[:topn]
    for k in sorted(ranked):
        r = data[k]
        print("  %5d %8.1fM %6.2f%%  %s" % (k, int(r[iex]) / 1e6, 100.0 * int(r[ismp]) / max(1, tots), r[isrc].strip()[:80]))
