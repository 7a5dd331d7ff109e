"""Split an ncu source-page CSV into phases at BAR.SYNC instructions and print instructions / samples per phase."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot_i = sum(int(r[col['Instructions Executed']] or 0) for r in body)
tot_s = sum(int(r[col['# Samples']] or 0) for r in body)
ph, acc_i, acc_s, n, first = 0, 0, 0, 0, None
def flush(tag):
    global acc_i, acc_s, n, ph
    print(f'phase {ph:2d} ({n:4d} SASS lines, ends at {tag[:40]:40s}): instr {acc_i / tot_i:6.1%}  samples {acc_s / tot_s:6.1%}')
    ph += 1; acc_i = acc_s = n = 0
for r in body:
    src = r[col['Source']].strip()
    acc_i += int(r[col['Instructions Executed']] or 0); acc_s += int(r[col['# Samples']] or 0); n += 1
    if 'BAR.SYNC' in src or 'EXIT' in src and n > 5:
        flush(src)
if n: flush('end')
