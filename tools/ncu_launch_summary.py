"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list as a markdown table.
    python tools/ncu_launch_summary.py launches.csv 'title' 'command' [raw list name] [kernel-name regex to keep] > profiles/xxx_summary.md"""
import collections, csv, re, sys
keep = re.compile(sys.argv[5]) if len(sys.argv) > 5 else None
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[col['Metric Name']] != 'gpu__time_duration.sum':
        continue
    v = float(r[col['Metric Value']].replace(',', '')); u = r[col['Metric Unit']]
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    name = r[col['Kernel Name']]
    if keep and not keep.search(name):
        continue
    short = name.split('(')[0].replace('void ', '').replace('fscnn::', '')
    a = agg.setdefault(short, [0, 0.0, r[col['Grid Size']], r[col['Block Size']]])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f'# {sys.argv[2]}\n\nCommand (B200, after the same command exited 0 without ncu):\n\n    {sys.argv[3]}\n')
print('Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolute times.')
print(f'Raw per-launch list: `{sys.argv[4] if len(sys.argv) > 4 else sys.argv[1]}`.  {sum(a[0] for a in agg.values())} launches, {tot:.0f} us in total.\n')
print('| kernel | launches | grid | block | total us | us / launch | share |\n|---|---|---|---|---|---|---|')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if a[1] / tot < 0.0005:
        continue
    print(f'| `{k[:80]}` | {a[0]} | {a[2]} | {a[3]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {a[1] / tot:.1%} |')
