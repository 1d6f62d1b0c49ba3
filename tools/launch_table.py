"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel markdown table."""
import collections, csv, sys

path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows[1:]:
    try:
        t = float(r[vi].replace(",", "")) / 1000.0
    except ValueError:
        continue
    name = r[ki].split("(")[0][:84]
    if "spin_kernel" in name:          # torch.cuda._sleep used by bench.py to keep the stream busy while it records events
        continue
    tot[name] = tot.get(name, 0.0) + t
    cnt[name] += 1
total = sum(tot.values())
print("| kernel | launches | total us | avg us | share |")
print("|---|---:|---:|---:|---:|")
for name in sorted(tot, key=lambda k: -tot[k]):
    print(f"| `{name}` | {cnt[name]} | {tot[name]:.1f} | {tot[name] / cnt[name]:.2f} | {100 * tot[name] / total:.1f}% |")
print(f"\n{sum(cnt.values())} launches, {total:.1f} us in total")
