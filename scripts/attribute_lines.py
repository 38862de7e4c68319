"""Attribute executed instructions of a profiled kernel to CUDA source lines.
usage: attribute_lines.py <nvdisasm --print-line-info dump> <ncu --page source --csv dump> <mangled kernel name>"""
import collections
import csv
import re
import sys

dump, srccsv, kname = sys.argv[1:4]
lines = open(dump).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.' + kname + ':')][0]
end = start + 1
while end < len(lines) and not lines[end].startswith('.text.'):
    end += 1
cur = None
instr = []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'^\s+/\*[0-9a-f]{4,5}\*/', l):
        instr.append((cur, l.strip()))
rows = list(csv.reader(open(srccsv)))
idx = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
startr = idx[-1]
hdr = rows[startr + 1]
ie, it, isamp = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
data = [(int(r[ie]), int(r[it]), int(r[isamp])) for r in rows[startr + 2:] if len(r) > it]
print("sass", len(instr), "profiled", len(data))
assert len(instr) == len(data)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for (loc, txt), (e, t, s) in zip(instr, data):
    a = agg[loc]
    a[0] += e; a[1] += t; a[2] += 1; a[3] += s
tot = sum(v[0] for v in agg.values())
tots = sum(v[3] for v in agg.values())
print("total warp instr", tot, "lanes", sum(v[1] for v in agg.values()) / tot)
srcs = {}
for loc, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    f, n = loc if loc else ("?", 0)
    try:
        if f not in srcs:
            srcs[f] = open('/root/repo/pysicalbasedraytracer_b200/csrc/' + f).read().split('\n')
        text = srcs[f][n - 1].strip()[:90]
    except Exception:
        text = ""
    print(f"{f}:{n:<4d} n={v[2]:4d} share={v[0]/tot:.3f} lanes={v[1]/max(v[0],1):5.1f} stall={v[3]/tots:.3f} | {text}")
