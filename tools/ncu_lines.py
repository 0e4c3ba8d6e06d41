#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (SASS view) per CUDA source line using nvdisasm -g output.
usage: ncu_lines.py <sass_page.csv> <nvdisasm_g.txt> <kernel-substring> <source.cu> [top]"""
import re, csv, sys
page, dis, kern, srcf = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
infunc = False; cur = None; addr2line = {}
for l in open(dis):
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m:
        infunc = kern in m.group(1); cur = None; continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
    if m and cur: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(page)))
hdr = rows[1]
si = hdr.index('# Samples'); ie = hdr.index('Instructions Executed'); ai = hdr.index('Address')
agg = {}; base = None
for r in rows[2:]:
    try: a = int(r[ai], 16)
    except Exception: continue
    if base is None: base = a
    key = addr2line.get(a - base, ('?', 0))
    x = agg.setdefault(key, [0, 0, 0]); x[0] += int(r[si] or 0); x[1] += int(r[ie] or 0); x[2] += 1
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print('samples', tot, 'warp-instructions', toti)
src = open(srcf).read().split('\n')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    line = src[k[1] - 1].strip()[:100] if (0 < k[1] <= len(src) and k[0] == srcf.split('/')[-1]) else ''
    print(f"{100*v[0]/tot:5.1f}%smp {100*v[1]/toti:5.1f}%ins sass={v[2]:4d} {k[0]}:{k[1]:5d}  {line}")
