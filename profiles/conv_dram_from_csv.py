"""Condense an ncu CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch of one
profiled step) into per-kernel-family totals and profiles/r2_conv_dram.json (DRAM bytes per trunk-conv launch,
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
layers = {"layer1": [0, 0.0], "layer2": [0, 0.0], "layer3": [0, 0.0]}


def layer_of(k):
    """trunk layer of a conv kernel instance, from its template arguments <cin, npad, ...> (TED generator)"""
    if k.startswith('conv128'):
        return "layer3"
    m = re.match(r'conv_tc_kernel<(\d+), (\d+)', k)
    if not m:
        return None
    cin, npad = int(m.group(1)), int(m.group(2))
    if cin == 32 and npad == 32:
        return "layer1"
    if npad == 64 and cin in (32, 64):
        return "layer2"
    return "layer3"          # 64 -> 128 (3x3 s2, 1x1 s2), 128 -> 34 final conv


for k, (n, us, rd, wr) in sorted(per.items(), key=lambda x: -x[1][1]):
    print(f"{k:48s} n={n:3d} {us:9.1f} us {100 * us / tot:5.1f}%  dram rd {rd / 1e6:9.1f} MB  wr {wr / 1e6:9.1f} MB  {(rd + wr) / us / 1e3 if us else 0:7.0f} GB/s")
    if k.startswith('conv'):
        conv_n += n
        conv_bytes += rd + wr
        lay = layer_of(k)
        if lay:
            layers[lay][0] += n
            layers[lay][1] += rd + wr
if len(sys.argv) > 3 and conv_n:
    json.dump({"clips": clips, "conv_launches": conv_n, "dram_bytes_per_launch": conv_bytes / conv_n,
               "layers": {k: {"launches": v[0], "dram_bytes_per_launch": v[1] / max(v[0], 1)} for k, v in layers.items()},
               "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, python profiles/step_once.py %d" % clips},
              open(sys.argv[3], 'w'), indent=1)
    print("wrote", sys.argv[3])
