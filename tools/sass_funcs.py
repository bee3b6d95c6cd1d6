#!/usr/bin/env python
"""Aggregate static / executed SASS by enclosing device function.
usage: sass_funcs.py <nvdisasm -g output> <kernel-substr> <ncu source csv> <physics.cuh> """
import csv, re, sys, collections
dis, fn, srccsv, cuh = sys.argv[1:5]
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and fn in l and l.endswith(":"))
addr_line = {}; cur = None
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith(".text."): break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m: addr_line[int(m.group(1), 16)] = (cur, m.group(2))
src = open(cuh).read().split("\n")
def fn_of(key):
    if key is None: return "?"
    f, line = key
    if f != "so101_physics.cuh": return f
    for i in range(line - 1, -1, -1):
        m = re.search(r'(?:SO101_DEV|__noinline__)\s+[\w<>]+\s+(\w+)\(', src[i])
        if m: return m.group(1)
    return "?"
rows = list(csv.reader(open(srccsv))); hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
stat = collections.Counter(); ex = collections.Counter(); th = collections.Counter(); smp = collections.Counter()
lineex = collections.Counter()
base = None
for r in rows[2:]:
    a = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
    if base is None: base = a
    key, _ = addr_line.get(a - base, (None, None))
    f = fn_of(key)
    stat[f] += 1
    e = int(r[ci["Instructions Executed"]] or 0)
    ex[f] += e; th[f] += int(r[ci["Thread Instructions Executed"]] or 0); smp[f] += int(r[ci["# Samples"]] or 0)
    lineex[key] += e
te, ts = sum(ex.values()) or 1, sum(smp.values()) or 1
print(f"{'function':28s} {'static':>7s} {'exec%':>7s} {'lanes':>6s} {'smpl%':>7s}")
for f, v in sorted(ex.items(), key=lambda kv: -kv[1]):
    print(f"{f:28s} {stat[f]:7d} {100*v/te:7.2f} {th[f]/max(1,v):6.1f} {100*smp[f]/ts:7.2f}")
if len(sys.argv) > 5:
    print("top lines by executed:")
    for k, v in lineex.most_common(int(sys.argv[5])):
        print(f"  {k}  {100*v/te:.2f}%  {src[k[1]-1].strip()[:90] if k and k[0]=='so101_physics.cuh' else ''}")
