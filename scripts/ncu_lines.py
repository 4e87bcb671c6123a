"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per source line."""
import csv, collections, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur_file = cur_fn = hdr = None
agg = collections.defaultdict(lambda: [0, 0])
with open(path) as f:
    for row in csv.reader(f):
        if not row: continue
        if row[0] == 'File Path': cur_file = row[1].split('/')[-1]; continue
        if row[0] == 'Function Name': cur_fn = row[1][:60]; continue
        if row[0] == 'Line No': hdr = row; continue
        if hdr is None or len(row) < 8 or row[2] != '-': continue
        d = dict(zip(hdr, row))
        try:
            ln = int(row[0]); ie = int(d['Instructions Executed']); sm = int(d['# Samples'])
        except Exception:
            continue
        k = (cur_fn, cur_file, ln, row[1].strip()[:100])
        agg[k][0] += ie; agg[k][1] += sm
for fn in sorted(set(k[0] for k in agg)):
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    tot = sum(v[0] for k, v in items); tots = sum(v[1] for k, v in items)
    print('=====', fn, 'inst', tot, 'samples', tots)
    for k, v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0]/max(tot,1)*100:5.1f}% inst {v[1]/max(tots,1)*100:5.1f}% smp  {k[1]}:{k[2]:4d}  {k[3]}")
