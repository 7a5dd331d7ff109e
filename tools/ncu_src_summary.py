"""Summarise an ncu `--page source --csv` dump: stall-reason totals, opcode mix, hottest SASS lines.
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv ; python tools/ncu_src_summary.py src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]
col_probe = hdr.index('# Samples')
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr) and r[0] != 'Address' and (r[col_probe] or '0').isdigit()]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter()
ops = collections.Counter()
opsamp = collections.Counter()
inst_total = 0
for r in body:
    for s in stalls:
        tot[s] += int(r[col[s]] or 0)
    op = r[col['Source']].split()
    op = op[1] if op and op[0].startswith('@') else (op[0] if op else '?')
    op = op.split('.')[0]
    n = int(r[col['Instructions Executed']] or 0)
    ops[op] += n
    opsamp[op] += int(r[col['# Samples']] or 0)
    inst_total += n
allsamp = sum(tot.values())
print('kernel rows', len(body), 'warp instructions', inst_total, 'samples', allsamp)
print('stall reasons:', ', '.join(f'{k[6:]} {v / allsamp:.1%}' for k, v in tot.most_common(8)))
print('opcode mix   :', ', '.join(f'{k} {v / inst_total:.1%}' for k, v in ops.most_common(14)))
print('samples by op:', ', '.join(f'{k} {v / max(1, sum(opsamp.values())):.1%}' for k, v in opsamp.most_common(10)))
top = sorted(body, key=lambda r: -int(r[col['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]
for r in top:
    rs = {s[6:]: int(r[col[s]] or 0) for s in stalls}
    main = max(rs, key=rs.get)
    print(f"  {int(r[col['# Samples']]):6d} samples  {main:12s} {r[col['Source']].strip()[:90]}")
