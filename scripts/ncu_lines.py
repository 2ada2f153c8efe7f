"""Instruction counts and stall samples per CUDA source line: joins `ncu --page source --csv` (SASS view) with
nvdisasm -g line info.  usage: ncu_lines.py <sass.csv> <nvdisasm.txt> <source file> [top]"""
import re, csv, sys
from collections import defaultdict
sass_csv, disasm, srcfile = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
cur = None; seq = []
for ln in open(disasm):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        off = int(m.group(1), 16)
        if seq and off == 0 and len(seq) > 10: break
        seq.append((off, cur, m.group(2)))
rows = list(csv.reader(open(sass_csv)))
h = rows[1]; ai = h.index('Address'); ie = h.index('Instructions Executed'); si = h.index('# Samples')
data = [(int(r[ai], 16), int(r[ie]), int(r[si])) for r in rows[2:] if len(r) > ie and r[ai].startswith('0x')]
base = data[0][0]
byoff = {a - base: (n, s) for a, n, s in data}
agg = defaultdict(lambda: [0, 0])
for off, c, txt in seq:
    if off in byoff:
        agg[c][0] += byoff[off][0]; agg[c][1] += byoff[off][1]
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total warp instructions", tot, "samples", ts)
base_name = srcfile.split('/')[-1]
src = open(srcfile).read().split('\n')
for (f, l), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src[l - 1].strip()[:100] if f == base_name else ''
    print(f"{n/tot*100:5.1f}% inst {s/ts*100:5.1f}% samp  {f}:{l}  {text}")
