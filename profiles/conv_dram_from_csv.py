"""Condense an ncu CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch of one
profiled step) into per-kernel-family totals and profiles/r1_conv_dram.json (DRAM bytes per trunk-conv launch,
read by bench.py for roofline.traffic).  usage: python profiles/conv_dram_from_csv.py launches.csv CLIPS [out.json]"""
import collections, csv, json, re, sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
clips = int(sys.argv[2])
per = collections.OrderedDict()
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3}
launch = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = re.sub(r'.*::', '', re.sub(r'\(.*', '', row['Kernel Name']))
    v = float(row['Metric Value'].replace(',', '')) * SCALE.get(row['Metric Unit'], 1)
    launch.setdefault((row['ID'], k), {})[row['Metric Name']] = v
for (_, k), m in launch.items():
    a = per.setdefault(k, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += m.get('gpu__time_duration.sum', 0.0)
    a[2] += m.get('dram__bytes_read.sum', 0.0)
    a[3] += m.get('dram__bytes_write.sum', 0.0)
tot = sum(a[1] for a in per.values())
print(f"clips {clips}  total {tot:.0f} us")
conv_n, conv_bytes = 0, 0.0
for k, (n, us, rd, wr) in sorted(per.items(), key=lambda x: -x[1][1]):
    print(f"{k:48s} n={n:3d} {us:9.1f} us {100 * us / tot:5.1f}%  dram rd {rd / 1e6:9.1f} MB  wr {wr / 1e6:9.1f} MB  {(rd + wr) / us / 1e3 if us else 0:7.0f} GB/s")
    if k.startswith('conv'):
        conv_n += n
        conv_bytes += rd + wr
if len(sys.argv) > 3 and conv_n:
    json.dump({"clips": clips, "conv_launches": conv_n, "dram_bytes_per_launch": conv_bytes / conv_n,
               "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, python profiles/step_once.py %d" % clips},
              open(sys.argv[3], 'w'), indent=1)
    print("wrote", sys.argv[3])
