"""Dev tool: one markdown table row per kernel of an .ncu-rep (raw page), plus the stall mix.
usage: ncu_summary.py <report.ncu-rep>"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, default=float('nan')):
    try:
        return float(r[col[name]].replace(',', ''))
    except Exception:
        return default


def to_bytes(r, name):
    v = val(r, name)
    u = units[col[name]] if name in col else ''
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)


def to_us(r, name):
    v = val(r, name)
    u = units[col[name]]
    return v * {'ns': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3}.get(u, 1)


print('| kernel | time us | dram read MB | dram write MB | dram % peak | fp64 pipe % | issue active % | warps active % | regs | smem excess wavefronts % |')
print('|---|---|---|---|---|---|---|---|---|---|')
for r in rows[2:]:
    name = r[col['Kernel Name']]
    rd, wr = to_bytes(r, 'dram__bytes_read.sum'), to_bytes(r, 'dram__bytes_write.sum')
    exc = val(r, 'derived__memory_l1_wavefronts_shared_excessive', 0.0)
    tot = val(r, 'smsp__inst_executed_op_shared_ld.sum', 0) + val(r, 'smsp__inst_executed_op_shared_st.sum', 0)
    print(f"| `{name}` | {to_us(r, 'gpu__time_duration.sum'):.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
          f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{val(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', val(r, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')):.1f} | "
          f"{val(r, 'sm__issue_active.avg.pct_of_peak_sustained_active', val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f} | "
          f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {int(val(r, 'launch__registers_per_thread', 0))} | "
          f"{exc:.0f} |")
    stalls = []
    for h in hdr:
        if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
            stalls.append((val(r, h, 0.0), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
    tot = sum(v for v, _ in stalls) or 1.0
    stalls.sort(reverse=True)
    print('\nstall mix `' + name.split('(')[0] + '`: ' + ', '.join(f'{n} {100 * v / tot:.0f}%' for v, n in stalls[:8]) + '\n')
    print(f'dram bytes per launch: {rd + wr:.0f}')
