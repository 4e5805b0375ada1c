"""ncu launch list (csv, --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum) -> per-kernel
averages per launch as JSON (profiles/r1_traffic.json; bench.py reads roofline.traffic from it).
usage: python tools/ncu_traffic.py launches.csv "workload text" > profiles/r1_traffic.json"""
import collections
import csv
import json
import sys

UNIT = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = next(r for r in rows if r[0] == "ID")
iK, iM, iU, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
per = collections.OrderedDict()
for r in rows:
    if not r[0].isdigit():
        continue
    k = r[iK].split("(")[0].replace("void ", "")
    d = per.setdefault(k, collections.defaultdict(list))
    d[r[iM]].append(float(r[iV].replace(",", "")) * UNIT.get(r[iU], 1.0))
out = {"workload": sys.argv[2] if len(sys.argv) > 2 else "", "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; "
       "averages per launch (per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes)", "kernels": {}}
tot = sum(sum(d["gpu__time_duration.sum"]) for d in per.values())
for k, d in per.items():
    n = len(d["gpu__time_duration.sum"])
    out["kernels"][k] = {"launches": n, "ms": sum(d["gpu__time_duration.sum"]) / n, "dram_read_bytes": sum(d["dram__bytes_read.sum"]) / n,
                         "dram_write_bytes": sum(d["dram__bytes_write.sum"]) / n, "share_of_profiled_time": sum(d["gpu__time_duration.sum"]) / tot}
print(json.dumps(out, indent=1))
