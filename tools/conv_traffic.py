"""Sum the DRAM traffic of the tcgen05 conv launches of ONE step from an ncu CSV
(`--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`) -> profiles/r2_conv_traffic.json."""
import collections, csv, json, re, sys
path, out = sys.argv[1], sys.argv[2]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
# group metric rows per launch ID
per = collections.OrderedDict()
for r in rows:
    per.setdefault(r['ID'], {'name': r['Kernel Name']})[r['Metric Name']] = (float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1, 'us': 1e3, 'ms': 1e6}
launches = []
for k, v in per.items():
    rd = v.get('dram__bytes_read.sum', (0, 'byte')); wr = v.get('dram__bytes_write.sum', (0, 'byte')); t = v.get('gpu__time_duration.sum', (0, 'ns'))
    launches.append((v['name'], rd[0] * scale.get(rd[1], 1), wr[0] * scale.get(wr[1], 1), t[0] * scale.get(t[1], 1)))
starts = [i for i, l in enumerate(launches) if 'clip_to_s2d' in l[0]]
a = starts[0] if starts else 0; b = starts[1] if len(starts) > 1 else len(launches)
step = [l for l in launches[a:b] if re.search(r'conv_umma|conv_halo|bottleneck_exit', l[0])]
res = {"conv_launches": len(step), "dram_read_bytes": int(sum(l[1] for l in step)), "dram_write_bytes": int(sum(l[2] for l in step)),
       "dram_bytes_per_step": int(sum(l[1] + l[2] for l in step)), "kernel_time_us": round(sum(l[3] for l in step) / 1e3, 1),
       "note": "ncu, cold caches per launch (cache control on), clocks uncontrolled; one eager step at B=8"}
if len(sys.argv) > 3:
    res["csrc_sha"] = sys.argv[3]     # hash of step_b200/csrc at capture time: bench.py only reports the traffic of the build it runs
json.dump(res, open(out, 'w'), indent=1)
print(res)
