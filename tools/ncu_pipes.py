"""Print per-kernel pipe utilisation (% of peak, active cycles) from an .ncu-rep: python tools/ncu_pipes.py rep [rep...]"""
import csv, subprocess, sys
WANT = ['alu', 'fma', 'fmaheavy', 'fmalite', 'lsu', 'xu', 'adu', 'uniform', 'cbu', 'tensor', 'fp16']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    kn = h.index('Kernel Name')
    def col(name):
        return h.index(name) if name in h else None
    cols = {p: col(f'sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active') for p in WANT}
    xu = [i for i, c in enumerate(h) if 'pipe_xu' in c]
    dur = col('gpu__time_duration.sum')
    issue = col('sm__inst_issued.avg.pct_of_peak_sustained_active') or col('smsp__issue_active.avg.pct')
    print(rep)
    print('kernel'.ljust(44), 'us'.rjust(7), ' '.join(p.rjust(8) for p in WANT), 'xu_rt'.rjust(7))
    for r in rows[2:]:
        vals = []
        for p in WANT:
            c = cols[p]
            vals.append(f'{float(r[c]):8.1f}' if c is not None and r[c] not in ('', 'n/a') else '       -')
        x = f'{float(r[xu[0]]):7.1f}' if xu and r[xu[0]] not in ('', 'n/a') else '      -'
        d = float(r[dur]) / 1e3 if dur is not None else 0
        print(r[kn][:44].ljust(44), f'{d:7.1f}', ' '.join(vals), x)
