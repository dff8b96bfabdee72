#!/usr/bin/env python
"""Top SASS instructions by stall samples for one kernel of an .ncu-rep (source page), plus an opcode histogram."""
import csv, io, subprocess, sys, collections
rep, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == 'Kernel Name':
        cur = {'name': row[1], 'hdr': None, 'rows': []}; blocks.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = row
    elif cur is not None and row:
        cur['rows'].append(row)
b = blocks[which]
h = b['hdr']
iS, iE, iT, iSamp = h.index('Source'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
tot = sum(int(r[iSamp]) for r in b['rows'])
totE = sum(int(r[iE]) for r in b['rows'])
print(b['name'][:80], 'samples', tot, 'warp-instr', totE, 'thread-instr', sum(int(r[iT]) for r in b['rows']))
rows = sorted(b['rows'], key=lambda r: -int(r[iSamp]))
for r in rows[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    top = sorted(((int(r[h.index(k)]), k) for k in stalls), reverse=True)[:2]
    print(f"{100*int(r[iSamp])/tot:5.1f}%  exec {int(r[iE]):>10}  thr/inst {int(r[iT])/max(1,int(r[iE])):5.1f}  {r[iS].strip()[:70]:70s} {top}")
hist = collections.Counter()
for r in b['rows']:
    op = r[iS].strip().split()[0] if r[iS].strip() else '?'
    if op.startswith('@'): op = r[iS].strip().split()[1]
    hist[op.split('.')[0]] += int(r[iE])
print('opcode histogram (warp instr):', [(k, f'{100*v/totE:.1f}%') for k, v in hist.most_common(25)])
agg = collections.Counter()
for r in b['rows']:
    for k in stalls: agg[k] += int(r[h.index(k)])
print('stall totals:', [(k, f'{100*v/tot:.1f}%') for k, v in agg.most_common(8)])
