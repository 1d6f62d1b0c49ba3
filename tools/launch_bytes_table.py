"""ncu CSV launch list with gpu__time_duration.sum + dram__bytes_{read,write}.sum -> per-kernel table (avg duration, DRAM
bytes per launch, achieved DRAM GB/s)."""
import collections, csv, sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ii, ki, mi, ui, vi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
per = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    except ValueError:
        continue
    d = per.setdefault(r[ii], {"name": r[ki].split("(")[0][:70]})
    d[r[mi]] = v
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d["name"], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
print("| kernel | launches | avg us | DRAM MB / launch | DRAM GB/s |")
print("|---|---:|---:|---:|---:|")
for name, (c, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {c} | {us / c:.1f} | {by / c / 1e6:.1f} | {by / us / 1e3 if us else 0:.0f} |")
