"""Bucket the warp-sampling counts of an `ncu --page source --csv --print-source sass` dump by SASS markers
(barriers, TMEM loads, global / shared accesses, calls): where the warps of a kernel spend their time."""
import csv, sys, itertools
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
def I(r, k):
    try: return int(r[ix[k]])
    except Exception: return 0
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
S = sum(I(r, '# Samples') for r in data)
print('samples', S, 'warp instructions', sum(I(r, 'Instructions Executed') for r in data))
tot = {s: sum(I(r, s) for r in data) for s in stalls}
print(' '.join(f"{s[6:]}={v / S:.3f}" for s, v in sorted(tot.items(), key=lambda x: -x[1])[:10]))
keys = sys.argv[2].split(',') if len(sys.argv) > 2 else ['BAR.SYNC', 'LDTM', 'UTCHMMA', 'CALL', 'UCGABAR', 'STG', 'ST.E', 'MEMBAR', 'ERRBAR', 'SYNCS', 'LDG', 'RET', 'LDL', 'STL']
acc = [I(r, '# Samples') for r in data]
pref = [0] + list(itertools.accumulate(acc))
prev = 0
for n, r in enumerate(data):
    s = r[ix['Source']]
    if any(t in s for t in keys):
        seg = pref[n] - pref[prev]
        if seg + acc[n] >= int(sys.argv[3]) if len(sys.argv) > 3 else True:
            top = {k[6:]: r[ix[k]] for k in stalls if r[ix[k]] not in ('0', '')}
            print(f"{r[ix['Address']][-5:]} n={n:5d} before={seg:6d} own={acc[n]:6d} exec={I(r, 'Instructions Executed'):9d} {s[:60]} {top if acc[n] > 300 else ''}")
        prev = n
