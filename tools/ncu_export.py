"""Export the judged numbers of .ncu-rep captures into a JSON summary:
   python tools/ncu_export.py out.json label=file.ncu-rep [label=file.ncu-rep ...]"""
import csv, json, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def export(path):
    out = open(path).read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        e = {"kernel": d['Kernel Name']}
        for k in KEYS:
            if k in d:
                e[k] = f"{d[k]} {u.get(k, '')}".strip()
        st = {}
        for k, v in d.items():
            if 'issue_stalled' in k and k.endswith('_per_issue_active.ratio'):
                try:
                    if float(v) >= 0.1:
                        st[k.split('issue_stalled_')[1].split('_per_issue')[0]] = round(float(v), 3)
                except ValueError:
                    pass
        e["stalls_per_issue"] = st
        res.append(e)
    return res


if __name__ == "__main__":
    out = {}
    for arg in sys.argv[2:]:
        label, path = arg.rsplit("=", 1)
        out[label] = export(path)
    json.dump(out, open(sys.argv[1], "w"), indent=1)
    print(json.dumps(out, indent=1)[:3000])
