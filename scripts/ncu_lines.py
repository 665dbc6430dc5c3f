#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g line info of the same cubin and
print executed warp-instructions (and stall samples) per CUDA source line.  usage: ncu_lines.py src.csv k.dis mangled_kernel [ticks]"""
import csv, re, sys
from collections import defaultdict
src, dis, kern = sys.argv[1:4]
ticks = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
# line info per instruction offset
off2line = {}
cur = None; inside = False
for l in open(dis):
    if l.startswith('.text.'):
        inside = l.startswith('.text.' + kern + ':')
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', l)
    if m: off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src)))
hdr = rows[1]; ia = hdr.index('Address'); ii = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); ithr = hdr.index('Thread Instructions Executed')
base = None
agg = defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[2:]:
    if len(r) <= ii: continue
    a = int(r[ia], 16)
    if base is None: base = a
    ln = off2line.get(a - base, ('?', 0))
    v = (int(r[ii]), int(r[isamp]), int(r[ithr]))
    for k in range(3): agg[ln][k] += v[k]; tot[k] += v[k]
print(f"total warp-instr {tot[0]:.3e}  per tick {tot[0]/ticks:.1f}  samples {tot[1]}  avg threads {tot[2]/max(tot[0],1):.1f}")
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 45]:
    print(f"{ln[0]}:{ln[1]:<5d} instr/tick {v[0]/ticks:7.2f}  ({100*v[0]/tot[0]:5.1f}%)  samples {100*v[1]/max(tot[1],1):5.1f}%  thr {v[2]/max(v[0],1):5.1f}")
