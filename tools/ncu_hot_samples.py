"""Like ncu_hot_lines.py but ranked by stall samples.
usage: ncu -i rep --page source --csv --print-source cuda,sass --kernel-name K | python tools/ncu_hot_samples.py [N]"""
import csv, sys
rows = list(csv.reader(sys.stdin)); n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
out = []; hdr = None
for r in rows:
    if len(r) > 8 and r[0] == 'Line No':
        hdr = r; sm = hdr.index('# Samples'); ie = hdr.index('Instructions Executed')
    elif hdr and len(r) > 8 and r[0].isdigit():
        try: out.append((int(r[sm]), int(r[ie]), int(r[0]), r[1].strip()[:100]))
        except ValueError: pass
ts = sum(o[0] for o in out) or 1
for o in sorted(out, reverse=True)[:n]: print(f'{o[0]/ts*100:5.1f}% smp | line {o[2]} | {o[3]}')
