"""Opcode histogram, stall reasons and the hottest SASS lines of one kernel from an .ncu-rep source page (read here).

    ncu -i x.ncu-rep --page source --csv --kernel-id :::1 > /tmp/src.csv ; python tools/ncu_sass_hot.py /tmp/src.csv
"""
import csv
import sys
from collections import Counter

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[1:]:
    if r == hdr:  # the export repeats the table for further launches: keep the first
        break
    if len(r) == len(hdr) and r[ix["# Samples"]].isdigit():
        data.append(r)
tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
totinst = sum(int(r[ix["Instructions Executed"]]) for r in data) or 1
print("samples", tot, "warp instructions", totinst, "SASS lines", len(data))
c, cs = Counter(), Counter()
for r in data:
    parts = r[ix["Source"]].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0]
    c[op] += int(r[ix["Instructions Executed"]])
    cs[op] += int(r[ix["# Samples"]])
for op, n in c.most_common(22):
    print("%-12s inst %10d (%4.1f%%)  samples %6d (%4.1f%%)" % (op, n, 100 * n / totinst, cs[op], 100 * cs[op] / tot))
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in data) for h in st}
print(", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print("%6s %9s  %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]][:100]))
