"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
   python scratch/launch_summary.py gpurun_out/r2_launches_c4.csv [skip_first_n]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
data = [r for r in rows if r is not hdr and r[ki] != "Kernel Name"][skip:]
agg = collections.OrderedDict()
for r in data:
    name = r[ki].split("(")[0].replace("odecol::tc::", "").replace("odecol::", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)     # -> microseconds
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
total = sum(a[1] for a in agg.values())
print(f"{len(data)} launches, {total / 1e3:.2f} ms of kernel time (serialised, cold-cache: compare shares)\n")
print("| kernel | launches | total us | mean us | share |")
print("|---|---:|---:|---:|---:|")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name[:90]}` | {n} | {t:.0f} | {t / n:.1f} | {100 * t / total:.1f} % |")
