"""Per-kernel markdown table from an `ncu --set full` report: time, DRAM traffic, occupancy, issue, pipes, smem, top stalls.
    python tools/ncu_kernel_table.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]; units = rows[1]
def col(n):
    return h.index(n) if n in h else None
def val(r, n, scale=1.0):
    c = col(n)
    if c is None or r[c] in ('', 'n/a', 'no data'):
        return float('nan')
    v = float(r[c].replace(',', ''))
    u = units[c]
    if u == 'Kbyte': v *= 1e3
    elif u == 'Mbyte': v *= 1e6
    elif u == 'Gbyte': v *= 1e9
    elif u in ('ns', 'nsecond'): v /= 1e3
    elif u in ('ms', 'msecond'): v *= 1e3
    return v * scale
stall_cols = [(i, c.split('issue_stalled_')[1].split('_per_issue')[0]) for i, c in enumerate(h)
              if c.startswith('smsp__average_warps_issue_stalled_') and c.endswith('_per_issue_active.ratio')]
print('| kernel | us | dram rd MB | dram wr MB | warps active % | issue active % | tensor pipe % | fma % | alu % | lsu % | smem LSU wavefronts % | smem tensor-core wavefronts % | warp instr (M) | regs | top stalls (warps per issue) |')
print('|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|')
for r in rows[2:]:
    name = r[col('Kernel Name')].replace('void ', '').replace('fscnn::', '').split('(')[0]
    st = []
    for i, nm in stall_cols:
        try: st.append((float(r[i]), nm))
        except ValueError: pass
    st = [s for s in sorted(st, reverse=True) if s[1] not in ('selected',)][:3]
    print(f"| `{name}` | {val(r, 'gpu__time_duration.sum'):.1f} | {val(r, 'dram__bytes_read.sum') / 1e6:.1f} | {val(r, 'dram__bytes_write.sum') / 1e6:.1f} "
          f"| {val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} "
          f"| {val(r, 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {val(r, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):.1f} "
          f"| {val(r, 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):.1f} | {val(r, 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):.1f} "
          f"| {val(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'):.1f} | {val(r, 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'):.1f} | {val(r, 'smsp__inst_executed.sum') / 1e6:.1f} "
          f"| {val(r, 'launch__registers_per_thread'):.0f} | {', '.join(f'{n} {v:.2f}' for v, n in st)} |")
