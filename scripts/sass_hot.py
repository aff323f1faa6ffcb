#!/usr/bin/env python
"""Per-region SASS execution profile of an .ncu-rep: python scripts/sass_hot.py rep tiles"""
import csv, subprocess, sys
rep, tiles = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()))
hdr = rows[1]
ia, isrc, ie, it, iss = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed'), hdr.index('# Samples')
out = []
for r in rows[2:]:
    try:
        e = float(r[ie])
    except Exception:
        continue
    out.append((r[ia], e / tiles, float(r[it] or 0), int(r[iss] or 0), r[isrc][:60]))
print("total per tile %.0f" % sum(o[1] for o in out))
seg, cur = [], None
for a, e, t, s, src in out:
    if cur and abs(cur['e'] - e) < 0.05 * max(cur['e'], 0.2):
        cur['n'] += 1; cur['samples'] += s; cur['last'] = src; cur['thr'] += t
    else:
        cur = {'e': e, 'n': 1, 'samples': s, 'first': src, 'last': src, 'addr': a, 'thr': t}; seg.append(cur)
tots = sum(c['samples'] for c in seg)
for c in seg:
    if c['e'] * c['n'] > 4:
        print("%s x%.2f n=%3d instr/tile=%4.0f thr=%2.0f samples=%4.1f%%  %s ... %s" % (c['addr'][-5:], c['e'], c['n'], c['e'] * c['n'], c['thr'] / c['n'], 100.0 * c['samples'] / tots, c['first'][:34], c['last'][:34]))
