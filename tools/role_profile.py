"""Per-role view of an `ncu --page source --csv` dump of a warp-specialised kernel (dev tool): prints every barrier wait,
TMEM load, bulk copy and MMA commit with its samples, so the waiting role shows up."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]
tot = sum(int(r[ix['# Samples']]) for r in data)
print(len(data), "lines,", tot, "samples,", sum(int(r[ix['Instructions Executed']]) for r in data), "warp instructions")
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 250
acc = 0
for i, r in enumerate(data):
    n = int(r[ix['# Samples']])
    acc += n
    if n >= thr or any(k in r[1] for k in ['TRYWAIT', 'UTCBAR', 'UBLKCP', 'BAR.SYNC', 'LDTM', 'EXIT']):
        st = {s: int(r[ix[s]] or 0) for s in stall}
        m = max(st, key=st.get)
        print(f"{i:5d} cum {acc:6d}  {n:6d} {r[ix['Instructions Executed']]:>9s}  {r[1].strip()[:64]:64s} {m} {st[m]}")
