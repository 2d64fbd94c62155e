"""Hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name K` (SASS view).
Usage: python scripts/ncu_hot.py source.csv [top_n]
Prints total samples / instructions, then the top instructions by stall samples with 2 lines of context."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], newline='')))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
head = rows[1]
ci = {h: i for i, h in enumerate(head)}
S, I, SRC = ci['# Samples'], ci['Instructions Executed'], ci['Source']
data = [r for r in rows[2:] if len(r) > max(S, I, SRC) and r[S].isdigit()]
tot_s = sum(int(r[S] or 0) for r in data)
tot_i = sum(int(r[I] or 0) for r in data)
print(f'{rows[0][1]}: samples {tot_s}, warp instructions {tot_i}, sass lines {len(data)}')
order = sorted(range(len(data)), key=lambda i: -int(data[i][S] or 0))[:top]
for i in order:
    r = data[i]
    print(f'{int(r[S]):7d} ({100.0 * int(r[S]) / max(tot_s, 1):5.1f}%) inst {int(r[I] or 0):10d}  [{i:4d}] {r[SRC].strip()[:110]}')
