"""Extract the judged metrics from `ncu -i X.ncu-rep --page raw --csv` output (stdin or file) into JSON."""
import csv, json, sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum"]
rows = list(csv.reader(open(sys.argv[1], errors="replace") if len(sys.argv) > 1 else sys.stdin))
hdr = next(r for r in rows if "Kernel Name" in r)
units = rows[rows.index(hdr) + 1]
out = {}
for r in rows[rows.index(hdr) + 2:]:
    if len(r) != len(hdr):
        continue
    name = r[hdr.index("Kernel Name")].split("(")[0]
    rec = {}
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            rec[k] = f"{r[i]} {units[i]}".strip()
    key, n = name, 1
    while key in out:
        n += 1
        key = f"{name} #{n}"
    out[key] = rec
print(json.dumps(out, indent=1))
