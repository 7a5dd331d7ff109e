"""Per-kernel warp stall breakdown (average warps stalled per issue-active cycle) from .ncu-rep files."""
import csv, subprocess, sys
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    kn = h.index('Kernel Name')
    idx = [(i, c.split('issue_stalled_')[1].split('_per_issue')[0]) for i, c in enumerate(h)
           if c.startswith('smsp__average_warps_issue_stalled_') and c.endswith('_per_issue_active.ratio')]
    print(rep)
    for r in rows[2:]:
        vals = []
        for i, nm in idx:
            try:
                vals.append((float(r[i]), nm))
            except ValueError:
                pass
        vals.sort(reverse=True)
        print(' ', r[kn][:48].ljust(48), '  '.join(f'{nm}={v:.2f}' for v, nm in vals[:7]))
