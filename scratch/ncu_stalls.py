"""Stall-reason totals and the hottest SASS instructions of one kernel from an ncu --set full report (source page):
   ncu -i rep --page source --csv --kernel-name regex:NAME --launch-count 1 > src.csv; python scratch/ncu_stalls.py src.csv [ntop]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[h]; data = [r for r in rows[h + 1:] if len(r) >= len(hdr) and r[0] != 'Address']
ix = {k: i for i, k in enumerate(hdr)}
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
I = lambda r, k: int(r[ix[k]] or 0)
ns = sum(I(r, '# Samples') for r in data)
ninst = sum(I(r, 'Instructions Executed') for r in data)
print('samples', ns, 'warp instructions', ninst, 'sass lines', len(data))
for s, v in sorted(((s, sum(I(r, s) for r in data)) for s in stalls), key=lambda x: -x[1])[:10]:
    print(f'  {s:24s} {v:8d} {100 * v / ns:5.1f} %')
print('hottest instructions (samples, executed, index, sass, top stall)')
for r in sorted(data, key=lambda r: -I(r, '# Samples'))[:ntop]:
    top = max(stalls, key=lambda s: I(r, s))
    print(f"{I(r, '# Samples'):6d} {I(r, 'Instructions Executed'):9d} {data.index(r):5d}  {r[ix['Source']].strip()[:80]:80s} {top}={I(r, top)}")
