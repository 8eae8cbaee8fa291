#!/usr/bin/env python3
"""Top SASS instructions by warp-stall samples: python tools_ncu_hot.py file.ncu-rep [N]"""
import csv, subprocess, sys
path = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(float(r[ci['Warp Stall Sampling (All Samples)']] or 0) for r in body)
print('total samples', tot, 'instructions', len(body))
order = sorted(range(len(body)), key=lambda i: -float(body[i][ci['Warp Stall Sampling (All Samples)']] or 0))
for i in order[:N]:
    r = body[i]
    print(f"{i:5d} {float(r[ci['Warp Stall Sampling (All Samples)']]) / tot * 100:6.2f}%  exec={r[ci['Instructions Executed']]:>10s}  {r[ci['Source']].strip()[:90]}")
