"""Per-block breakdown of an ncu launch list (gpu__time_duration.sum CSV) of tools/prof_iter.py."""
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
idx = [i for i, x in enumerate(rows) if 'k_stem' in x['Kernel Name']]
seq = rows[idx[-1]:idx[-1] + 110]
names = []
for x in seq:
    n, t = x['Kernel Name'], float(x['Metric Value']) / 1e3
    short = ('stem' if 'k_stem' in n else 'dw' if 'dwconv' in n else 'se' if 'se_gate' in n else
             'gemm' if 'gemm' in n else 'pool' if 'pool' in n else n[:10])
    names.append((short, t, x['Grid Size']))
b, j, te, tp, td, ts = 0, 1, 0, 0, 0, 0
print('stem', names[0][1])
while j < len(names):
    if names[j][0] == 'pool':
        print('pool', names[j][1])
        break
    if names[j][0] == 'gemm' and names[j + 1][0] == 'dw':
        e, d, se, p = names[j:j + 4]
        j += 4
        te += e[1]; tp += p[1]; td += d[1]; ts += se[1]
        print(f'block{b:2d} expand {e[1]:7.1f} {e[2]:>12s} dw {d[1]:7.1f} se {se[1]:5.1f} proj {p[1]:7.1f} {p[2]:>12s}')
    elif names[j][0] == 'dw':
        d, se, p = names[j:j + 3]
        j += 3
        tp += p[1]; td += d[1]; ts += se[1]
        print(f'block{b:2d} expand    -                 dw {d[1]:7.1f} se {se[1]:5.1f} proj {p[1]:7.1f} {p[2]:>12s}')
    elif names[j][0] == 'gemm':
        print('head gemm', names[j][1], names[j][2])
        j += 1
        continue
    b += 1
print('total us', round(sum(t for _, t, _ in names), 1), 'expand', round(te, 1), 'proj', round(tp, 1), 'dw', round(td, 1), 'se', round(ts, 1))
