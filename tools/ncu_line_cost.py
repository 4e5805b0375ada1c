"""Per-source-line instruction cost of one kernel from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu -i rep --page source --csv --print-source cuda,sass --kernel-name K | python tools/ncu_line_cost.py UNITS [MIN]
Prints thread-instructions per UNIT (e.g. per output line / per array element) for every line above MIN."""
import csv
import sys

units = float(sys.argv[1])
mn = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = list(csv.reader(sys.stdin))
out, fname, hdr = [], "", None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
        ie, te, sm = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    elif hdr and len(r) > 8 and r[0].isdigit():
        try:
            out.append((fname, int(r[0]), int(r[ie]), int(r[sm]), r[1].strip()))
        except ValueError:
            pass
tot = sum(o[2] for o in out)
ts = sum(o[3] for o in out) or 1
print(f"warp-inst {tot}  -> {tot * 32 / units:.1f} thread-inst per unit")
for o in sorted(out, key=lambda o: (o[0], o[1])):
    if o[2] * 32 / units >= mn or o[3] / ts > 0.01:
        print(f"{o[0]}:{o[1]:5d} {o[2] * 32 / units:7.1f}/unit smp {o[3] / ts * 100:5.1f}% | {o[4][:120]}")
