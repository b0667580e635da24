#!/usr/bin/env python
"""profiles/traffic_<workload>.json from an `ncu --set full` raw page: DRAM bytes per launch of the captured kernel.
usage: traffic_from_ncu.py <raw.csv> <out.json> <points per launch>
bench.py only attaches the figure to a roofline line when the kernel it launched carries the same name."""
import csv
import json
import sys

raw, out, points = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot, names = [], set()
for r in rows[2:]:
    b = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        b += float(r[col[m]]) * scale[units[col[m]]]
    tot.append(b)
    names.add(r[col["Kernel Name"]])
assert len(names) == 1, names
json.dump({"kernel": names.pop(), "dram_bytes_per_launch": sum(tot) / len(tot), "points": points, "source": "profiles/" + raw.split("/")[-1],
           "note": "dram__bytes_read.sum + dram__bytes_write.sum, mean over captured launches (ncu --set full); per launch of `points` points"},
          open(out, "w"))
print(open(out).read())
