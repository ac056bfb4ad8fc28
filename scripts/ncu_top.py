"""Print the hottest SASS lines (stall samples) of kernel #k from an `ncu --page source --csv` export."""
import csv, sys
path, kid = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
kern, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}; kern.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr'] = r; continue
    cur['rows'].append(r)
k = kern[kid]; h = k['hdr']
si, ii, so = h.index('# Samples'), h.index('Instructions Executed'), h.index('Source')
stall = [i for i, nme in enumerate(h) if nme.startswith('stall_')]
tot = sum(int(r[si] or 0) for r in k['rows'])
print(len(kern), 'kernels; kernel', kid, k['name'][:60], 'total samples', tot)
agg = {}
for r in k['rows']:
    for i in stall:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
print('stall totals:', sorted(agg.items(), key=lambda x: -x[1])[:8])
for r in sorted(k['rows'], key=lambda r: -int(r[si] or 0))[:n]:
    st = sorted([(int(r[i] or 0), h[i][6:]) for i in stall], reverse=True)[:2]
    print(r[si].rjust(6), r[ii].rjust(8), r[so][:100].ljust(100), st)
