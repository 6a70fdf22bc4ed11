"""Dev tool: top stall locations of an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv > file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print('total samples', tot)
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else len(data)
if len(sys.argv) > 3:
    for i in range(lo, hi):
        r = data[i]
        s = int(r[ix['# Samples']] or 0)
        top = sorted(((int(r[ix[k]] or 0), k) for k in stalls), reverse=True)[:2]
        print(i, s, r[ix['Instructions Executed']], r[ix['Source']][:90], [(k[6:], v) for v, k in top if v])
else:
    order = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']] or 0))[:n]
    for i in order:
        r = data[i]
        s = int(r[ix['# Samples']] or 0)
        top = sorted(((int(r[ix[k]] or 0), k) for k in stalls), reverse=True)[:3]
        print(i, s, '%.1f%%' % (100 * s / tot), r[ix['Source']][:80], [(k[6:], v) for v, k in top if v])
