"""Per-source-line instruction and stall-sample shares from an ncu report.
    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > x.csv ; python tools/ncu_lines.py x.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cur = None; hdr = None; out = []
for r in rows:
    if r and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; hdr = None; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == '-':   # a source line row (aggregated over its SASS)
        i_ie = hdr.index('Instructions Executed'); i_s = hdr.index('# Samples')
        try:
            ie = float(r[i_ie] or 0); sm = float(r[i_s] or 0)
        except ValueError:
            continue
        if ie > 0 or sm > 0:
            out.append((ie, sm, cur, r[0], r[1].strip()[:100]))
ti = sum(o[0] for o in out); ts = sum(o[1] for o in out)
print('warp instructions', ti, 'samples', ts)
for o in sorted(out, reverse=True)[:top]:
    print(f"{o[0]/ti*100:5.1f}% inst {o[1]/max(ts,1)*100:5.1f}% smpl  {o[2]}:{o[3]}  {o[4]}")
