"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the kernels of the last step, in launch order (dev tool)."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[1:]
names = [r[ix['Kernel Name']] for r in data]
casc = [i for i, n in enumerate(names) if 'cascade' in n]
first = casc[-1]
while first > 0 and 'cascade' in names[first - 1]:
    first -= 1
tot = 0.0
for r in data[first:]:
    t = float(r[ix['Metric Value']].replace(',', ''))
    tot += t
    print(f"{t / 1000:9.1f} us  {r[ix['Kernel Name']][:80]}  grid {r[ix['Grid Size']]} block {r[ix['Block Size']]}")
print(f"{tot / 1e6:.3f} ms in {len(data) - first} launches")
