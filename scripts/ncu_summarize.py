"""Turns `ncu -i X.ncu-rep --page raw --csv` output into the compact JSON kept under profiles/.

Usage: python scripts/ncu_summarize.py raw.csv out.json [skip_launches]

Every row is one kernel launch; values are normalised to bytes / microseconds / percent using the
unit row ncu prints under the header.  Nothing here runs on the GPU."""
import csv
import json
import sys

KEEP = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
]
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12,
         'ns': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = list(csv.reader(open(src, newline='')))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    head, units = rows[hi], rows[hi + 1]
    col = {h: i for i, h in enumerate(head)}
    out = []
    for r in rows[hi + 2 + skip:]:
        if len(r) != len(head):
            continue
        d = {'kernel': r[col['Kernel Name']]}
        for k in KEEP:
            if k not in col:
                continue
            try:
                v = float(r[col[k]].replace(',', ''))
            except ValueError:
                continue
            u = units[col[k]]
            d[k] = v * SCALE.get(u, 1)
            if u in SCALE:
                d[k + '.unit'] = 'us' if 'time' in k else 'byte'
        out.append(d)
    json.dump(out, open(dst, 'w'), indent=1)
    for d in out:
        t = d.get('gpu__time_duration.sum', 0.0)
        print(f"{d['kernel'][:70]:70s} {t:9.1f} us  rd {d.get('dram__bytes_read.sum', 0) / 1e6:9.1f} MB  "
              f"wr {d.get('dram__bytes_write.sum', 0) / 1e6:9.1f} MB  "
              f"tensor {d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):5.1f}%")


if __name__ == '__main__':
    main()
