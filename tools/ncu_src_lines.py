"""Aggregate an ncu `--page source --csv --print-source=cuda,sass` dump per CUDA source line:
    ncu -i rep --page source --csv --print-source=cuda,sass --kernel-name regex:NAME > src.csv; python tools/ncu_src_lines.py src.csv [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = '?'
agg = collections.defaultdict(lambda: [0, 0, 0, ''])   # instr, samples, smem wavefronts, text
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        hdr = r; col = {}
        for i, h in enumerate(r):
            col.setdefault(h, i)
        continue
    if hdr is None or len(r) != len(hdr) or not r[0].isdigit():
        continue
    k = (cur, int(r[0]))
    a = agg[k]
    def num(name):
        i = col.get(name)
        try: return int(float(r[i])) if i is not None and r[i] else 0
        except ValueError: return 0
    a[0] += num('Instructions Executed'); a[1] += num('# Samples')
    a[2] += num('L1 Wavefronts Shared') or num('L1 Wavefronts Shared Excessive') * 0
    a[3] = r[1].strip()[:110]
ti = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
print(f'total warp instr {ti}, samples {ts}')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{k[0][:18]:18s}:{k[1]:4d} instr {a[0] / ti:6.1%} samples {a[1] / ts:6.1%} smemwf {a[2]:9d} | {a[3]}')
