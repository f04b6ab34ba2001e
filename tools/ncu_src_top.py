"""Summarise an `ncu --page source --csv` dump: stall totals and the hottest SASS lines (dev tool)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]
tot = sum(int(r[ix['# Samples']]) for r in data)
print(len(data), "sass lines; samples", tot, "; warp instructions", sum(int(r[ix['Instructions Executed']]) for r in data))
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stall}
print(sorted(agg.items(), key=lambda x: -x[1])[:10])
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:n_top]:
    st = {s: int(r[ix[s]] or 0) for s in stall}
    m = max(st, key=st.get)
    print(r[ix['# Samples']], r[ix['Instructions Executed']], r[1].strip()[:80], m, st[m])
