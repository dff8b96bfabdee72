import csv, sys, collections, re
tot = collections.defaultdict(float); cnt = collections.Counter()
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hdr = None
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    name = r[hdr.index('Kernel Name')]; v = r[hdr.index('Metric Value')]; u = r[hdr.index('Metric Unit')]
    try: t = float(v.replace(',', ''))
    except ValueError: continue
    t *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}.get(u, 1e-6)
    k = re.sub(r'<.*', '', name.split('(')[0]).split('::')[-1]
    tot[k] += t; cnt[k] += 1
for k, t in sorted(tot.items(), key=lambda kv: -kv[1]): print(f"{t:9.3f} ms  {cnt[k]:4d}  {k}")
print(f"{sum(tot.values()):9.3f} ms total")
