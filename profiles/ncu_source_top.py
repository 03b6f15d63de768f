"""Top stall-sample instructions of a kernel from an .ncu-rep source page (SASS view).
usage: python profiles/ncu_source_top.py rep.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
lines = out.splitlines()
rows = list(csv.reader(lines[1:]))
h = rows[0]
ci = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
data = []
for idx, r in enumerate(rows[1:]):
    try:
        s = int(r[ci['# Samples']])
    except Exception:
        continue
    data.append((s, idx, r))
tot = sum(d[0] for d in data)
print('total samples', tot)
for s, idx, r in sorted(data, reverse=True)[:N]:
    top = sorted(((int(r[ci[n]] or 0), n) for n in stalls), reverse=True)[:3]
    print(f"{s:7d} {100.0*s/tot:5.1f}%  #{idx:5d} {r[ci['Source']].strip()[:70]:70s} " + ' '.join(f"{n[6:]}={v}" for v, n in top if v))
