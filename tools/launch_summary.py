"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: one step (s2d -> next s2d)."""
import collections, csv, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
names = [(re.sub(r'\(.*', '', x['Kernel Name']).replace('void ', '').replace('step::', ''), float(x['Metric Value']), x['Grid Size']) for x in rows]
starts = [i for i, (n, t, g) in enumerate(names) if 'clip_to_s2d' in n]
a = starts[0]; b = starts[1] if len(starts) > 1 else len(names)
step = names[a:b]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, t, g in step:
    agg[n][0] += 1; agg[n][1] += t
tot = sum(v[1] for v in agg.values())
print("launches in one step: %d, summed kernel time: %.3f ms" % (len(step), tot / 1e6))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-42s %4d launches %9.1f us %5.1f%%" % (n[:42], c, t / 1e3, 100 * t / tot))
if len(sys.argv) > 2:
    print()
    for i, (n, t, g) in enumerate(step):
        print("%3d %-40s %9.1f us grid %s" % (i, n[:40], t / 1e3, g))
