"""Aggregate `ncu --page source --csv --print-source cuda,sass` by CUDA source line.
usage: ncu -i rep --page source --csv --print-source cuda,sass --kernel-name K | python tools/ncu_hot_lines.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
n_top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
out, fname, hdr = [], "", None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
        ie, te, sm = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    elif hdr and len(r) > 8 and r[0].isdigit():
        try:
            out.append((int(r[ie]), int(r[te]), int(r[sm]), fname, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot = sum(o[0] for o in out) or 1
ts = sum(o[2] for o in out) or 1
print(f"total warp instructions {tot}, samples {ts}")
for o in sorted(out, reverse=True)[:n_top]:
    print(f"{o[0]/tot*100:5.1f}% inst {o[2]/ts*100:5.1f}% smp thr/inst {o[1]/max(o[0],1):5.1f} | {o[3]}:{o[4]} | {o[5][:100]}")
