"""dev tool: markdown table of one `ncu --set full` report (`ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/ncu_table.py raw.csv`)."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name, unit_to=None):
    i = ix[name]
    v = float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else 0.0
    u = units[i]
    if unit_to == "MB":
        v *= {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
    if unit_to == "us":
        v *= {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
    return v


print("| kernel | time us | DRAM rd MB | DRAM wr MB | L2 MB | SM % | tensor pipe % | issue % | warps % | smem wavefronts % | regs | grid | block |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
tot = 0.0
for r in data:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").strip()
    t = val(r, "gpu__time_duration.sum", "us")
    tot += t
    print("| `%s` | %.1f | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %s | %s |" % (
        name, t, val(r, "dram__bytes_read.sum", "MB"), val(r, "dram__bytes_write.sum", "MB"), val(r, "lts__t_sectors.sum") * 32 / 1e6,
        val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"), val(r, "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        val(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"), val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"), int(val(r, "launch__registers_per_thread")),
        r[ix["Grid Size"]] if "Grid Size" in ix else "", r[ix["Block Size"]] if "Block Size" in ix else ""))
print("\nsum of the launches: %.1f us" % tot)
