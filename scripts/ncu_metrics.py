"""Print selected metrics of kernels from an `ncu --page raw --csv` dump.
Usage: python scripts/ncu_metrics.py raw.csv <kernel substring>[:nth] [metric substrings...]"""
import csv
import sys

DEFAULT = ['gpu__time_duration.sum', 'tc_wavefronts_mem_shared.sum.pct', 'lsu_wavefronts_mem_shared.sum.pct',
           'lsu_wavefronts_mem_shared_op_ld.sum.pct', 'lsu_wavefronts_mem_shared_op_st.sum.pct',
           'l1tex__data_pipe_lsu_wavefronts.sum.pct', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct',
           'dram__throughput.avg.pct', 'lts__throughput.avg.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
           'issue_stalled', 'sm__inst_executed_pipe_tmem.avg', 'sm__warps_active.avg.pct', 'registers_per_thread',
           'l1tex__m_xbar2l1tex_read_bytes.sum ', 'l1tex__m_l1tex2xbar_write_bytes.sum ', 'smsp__inst_executed.sum ']

rows = list(csv.reader(open(sys.argv[1], newline='')))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
head, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
kn = head.index('Kernel Name')
sel = sys.argv[2].split(':')
match = [r for r in data if sel[0] in r[kn]]
r = match[int(sel[1]) if len(sel) > 1 else -1]
subs = sys.argv[3:] or DEFAULT
print(r[kn])
for i, h in enumerate(head):
    if any(s.strip() in h and (not s.endswith(' ') or h.endswith(s.strip())) for s in subs):
        try:
            v = float(r[i].replace(',', ''))
        except ValueError:
            continue
        if v != 0:
            print(f'  {h[:105]:105s} {units[i]:10s} {r[i]}')
