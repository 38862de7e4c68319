"""Attribute the executed instructions of a profiled kernel to the OUTERMOST source line inside the kernel body (inlined callees
are charged to their call site) - i.e. to the phases of the kernel's main loop.
usage: attribute_outer.py <nvdisasm --print-line-info-inline dump> <ncu --page source --csv dump> <mangled kernel name> <kernel source file>"""
import collections
import csv
import re
import sys

dump, srccsv, kname, kfile = sys.argv[1:5]
lines = open(dump).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.' + kname + ':')][0]
end = start + 1
while end < len(lines) and not lines[end].startswith('.text.'):
    end += 1
instr = []
cur = None
pending = []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)( inlined at)?', l)
    if m:
        pending.append((m.group(1).split('/')[-1], int(m.group(2)), bool(m.group(3))))
        continue
    if re.match(r'^\s+/\*[0-9a-f]{4,5}\*/', l):
        if pending:
            outer = [p for p in pending if not p[2]]
            cur = (outer[-1][0], outer[-1][1]) if outer else (pending[-1][0], pending[-1][1])
            inner = (pending[0][0], pending[0][1])
            pending = []
            curinner = inner
        instr.append((cur, curinner if cur else None, l.strip()))
rows = list(csv.reader(open(srccsv)))
idx = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
startr = idx[-1]
hdr = rows[startr + 1]
ie, it, isamp = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
data = [(int(r[ie]), int(r[it]), int(r[isamp])) for r in rows[startr + 2:] if len(r) > it]
assert len(instr) == len(data), (len(instr), len(data))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for (loc, inner, txt), (e, t, s) in zip(instr, data):
    a = agg[loc]
    a[0] += e; a[1] += t; a[2] += 1; a[3] += s
tot = sum(v[0] for v in agg.values())
tots = sum(v[3] for v in agg.values())
print("total warp instr", tot, "lanes", sum(v[1] for v in agg.values()) / tot)
src = open('/root/repo/pysicalbasedraytracer_b200/csrc/' + kfile).read().split('\n')
for loc, v in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    f, n = loc if loc else ("?", 0)
    text = src[n - 1].strip()[:100] if f == kfile and 0 < n <= len(src) else ""
    if v[0] / tot >= 0.002:
        print(f"{f}:{n:<4d} sass={v[2]:4d} share={v[0]/tot:.3f} lanes={v[1]/max(v[0],1):5.1f} stall={v[3]/max(tots,1):.3f} | {text}")
