"""brief per-kernel summary of an .ncu-rep: python tools/ncu_brief.py file.ncu-rep"""
import csv, subprocess, sys
out = open(sys.argv[1]).read() if sys.argv[1].endswith(".csv") else subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__inst_executed_pipe_lsu.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'lts__t_sectors_srcunit_tex_op_read.sum']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('=====', d['Kernel Name'])
    for k in want:
        if k in d: print('  ', k, d[k])
    for k, v in d.items():
        if 'issue_stalled' in k and k.endswith('_per_issue_active.ratio'):
            try:
                if float(v) >= 0.1: print('   stall', k.split('issue_stalled_')[1].split('_per_issue')[0], v)
            except ValueError:
                pass
