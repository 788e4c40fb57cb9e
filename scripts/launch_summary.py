"""Summarise an ncu --metrics gpu__time_duration.sum launch list: per-kernel totals for ONE train
step (the launches between the last two adam_kernel launches of the resident loop) and the slowest
individual launches."""
import csv, collections, sys
path = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 150.0
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
def us(r):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    return v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
adam = [i for i, r in enumerate(rows) if r['Kernel Name'].startswith(('adam_kernel', 'adam_dev_kernel'))]
s, e = adam[-2] + 1, adam[-1] + 1          # the last complete step
step = rows[s:e]
tot = sum(us(r) for r in step)
print(f"launches in step: {len(step)}   sum of kernel time: {tot/1e3:.2f} ms")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in step:
    a = agg[r['Kernel Name'].split('(')[0][:70]]; a[0] += 1; a[1] += us(r)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{v[1]/1e3:8.3f} ms {100*v[1]/tot:5.1f}%  n={v[0]:4d}  {k}")
print("--- slowest launches")
for i, r in enumerate(step):
    t = us(r)
    if t > thr:
        print(f"{i:4d} {t:9.1f} us grid={r['Grid Size']:>14s} {r['Kernel Name'].split('(')[0][:60]}")
