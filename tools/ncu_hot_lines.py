#!/usr/bin/env python
"""Per CUDA source line: stall samples, warp instructions, threads per instruction, for kernel #k of an .ncu-rep."""
import csv, io, subprocess, sys, collections, os
rep, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
kern_order, data = [], collections.OrderedDict()
f = fn = hdr = None
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == 'File Path': f = row[1]; continue
    if row[0] == 'Function Name':
        fn = row[1]
        if fn not in kern_order: kern_order.append(fn)
        continue
    if row[0] == 'Line No': hdr = row; continue
    if row[0] == '' or hdr is None: continue  # sass rows
    try: line = int(row[0])
    except ValueError: continue
    iS, iE, iT = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
    key = (fn, os.path.basename(f), line)
    d = data.setdefault(key, [0, 0, 0, row[1]])
    g = lambda s: int(s) if s.lstrip("-").isdigit() else 0
    d[0] += g(row[iS]); d[1] += g(row[iE]); d[2] += g(row[iT])
fnsel = kern_order[which]
rows = [(k, v) for k, v in data.items() if k[0] == fnsel]
totS = sum(v[0] for _, v in rows); totE = sum(v[1] for _, v in rows)
print(fnsel[:90], 'samples', totS, 'warp-instr', totE)
for k, v in sorted(rows, key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*v[1]/totE:5.1f}% instr {100*v[0]/max(totS,1):5.1f}% stall  thr/inst {v[2]/max(v[1],1):5.1f}  {k[1]}:{k[2]:<4d} {v[3].strip()[:95]}")
