"""Key metrics per launch from an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = [('Kernel Name', 'kernel'), ('launch__grid_size', 'grid'), ('launch__block_size', 'blk'), ('gpu__time_duration.sum', 'us'),
        ('dram__bytes_read.sum', 'rdMB'), ('dram__bytes_write.sum', 'wrMB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
        ('launch__registers_per_thread', 'regs'), ('launch__occupancy_limit_shared_mem', 'lim_smem'),
        ('launch__occupancy_limit_registers', 'lim_reg'), ('smsp__inst_executed.sum', 'winst')]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
tens = [i for i, h in enumerate(hdr) if 'pipe_tensor' in h and 'pct' in h]
print(' '.join('%s' % n for _, n in idx))
for r in rows[2:]:
    out = []
    for i, n in idx:
        v = r[i]
        if n == 'kernel': v = v.split('(')[0].split('::')[-1][:28]
        else:
            try: v = '%.4g' % float(v.replace(',', ''))
            except ValueError: pass
        out.append(v)
    print(' '.join(out), '| tensor:', ' '.join(r[i] for i in tens[:3]))
