"""Dump instructions [a,b) of the SASS source page with samples and top stall reason."""
import csv, subprocess, sys
rep = sys.argv[1]; a = int(sys.argv[2]); b = int(sys.argv[3])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()[1:]))
h = rows[0]; ci = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
for idx, r in enumerate(rows[1:]):
    if idx < a or idx >= b: continue
    s = int(r[ci['# Samples']] or 0)
    top = sorted(((int(r[ci[n]] or 0), n) for n in stalls), reverse=True)[:2]
    print(f"#{idx:5d} {s:6d} ex={r[ci['Instructions Executed']]:>8s} {r[ci['Source']].strip()[:90]:90s} " + ' '.join(f"{n[6:]}={v}" for v, n in top if v))
