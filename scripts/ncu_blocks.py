"""Aggregate an `ncu --page source --csv --print-source sass` dump into runs of instructions with (nearly) equal execution counts
(= basic blocks / loop bodies): address range, instructions, executions, share of instructions and of stall samples, opcode mix.
usage: ncu_blocks.py dump.csv [min_share_percent]"""
import csv, collections, sys
path = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = []; hdr = None
with open(path) as f:
    for row in csv.reader(f):
        if not row: continue
        if row[0] == 'Address': hdr = row; continue
        if hdr and row[0].startswith('0x'):
            d = dict(zip(hdr, row))
            rows.append((int(row[0], 16), row[1].strip(), int(d['Instructions Executed']), int(d['# Samples'])))
base = rows[0][0]
tot = sum(r[2] for r in rows); ts = sum(r[3] for r in rows)
print('instructions executed', tot, 'samples', ts)
def op(s):
    p = s.split()
    return (p[1] if p[0].startswith('@') else p[0]).split('.')[0]
blocks = []; cur = None
for a, s, ie, sm in rows:
    if cur and abs(ie - cur['ie']) <= 0.03 * max(cur['ie'], 1):
        cur['n'] += 1; cur['tot'] += ie; cur['sm'] += sm; cur['end'] = a; cur['ops'].append(op(s))
    else:
        if cur: blocks.append(cur)
        cur = {'start': a, 'end': a, 'ie': ie, 'n': 1, 'tot': ie, 'sm': sm, 'ops': [op(s)]}
blocks.append(cur)
for b in blocks:
    if b['tot'] / tot * 100 > thr:
        c = collections.Counter(b['ops'])
        print(f"{b['start']-base:6x}-{b['end']-base:6x} n={b['n']:4d} exec={b['ie']/1e6:8.2f}M {b['tot']/tot*100:5.1f}% inst {b['sm']/max(ts,1)*100:5.1f}% smp  {dict(c.most_common(7))}")
