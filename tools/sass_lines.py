#!/usr/bin/env python
"""Join `nvdisasm -g` line info with an ncu source-page CSV: per source line static SASS count,
executed warp instructions and stall samples.  usage: sass_lines.py <dis.txt> <func-substr> [src.csv]"""
import csv, re, sys, collections
dis, fn = sys.argv[1], sys.argv[2]
srccsv = sys.argv[3] if len(sys.argv) > 3 else None
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and fn in l and l.endswith(":"))
addr_line = {}
cur = None
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith(".text."):
        break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        addr_line[int(m.group(1), 16)] = (cur, m.group(2))
static = collections.Counter(v[0] for v in addr_line.values())
execd = collections.Counter(); samples = collections.Counter(); noinst = collections.Counter()
if srccsv:
    rows = list(csv.reader(open(srccsv)))
    hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
    base = None
    for r in rows[2:]:
        a = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
        if base is None: base = a
        key = addr_line.get(a - base, (None,))[0]
        execd[key] += int(r[ci["Instructions Executed"]] or 0)
        samples[key] += int(r[ci["# Samples"]] or 0)
        noinst[key] += int(r[ci["stall_no_inst"]] or 0)
tot_s, tot_e, tot_p = sum(static.values()), sum(execd.values()) or 1, sum(samples.values()) or 1
print(f"total static {tot_s}  executed {tot_e}  samples {tot_p}")
print(f"{'file:line':40s} {'static':>7s} {'exec%':>7s} {'smpl%':>7s} {'noinst%':>8s}")
for k, v in sorted(static.items(), key=lambda kv: -kv[1])[: int(sys.argv[4]) if len(sys.argv) > 4 else 60]:
    name = f"{k[0]}:{k[1]}" if k else "?"
    print(f"{name:40s} {v:7d} {100*execd[k]/tot_e:7.2f} {100*samples[k]/tot_p:7.2f} {100*noinst[k]/max(1,samples[k]):8.1f}")
