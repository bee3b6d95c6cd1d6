#!/usr/bin/env python
"""Per-role view of an ncu source-page CSV of a team (SPLIT) kernel.
usage: team_roles.py <nvdisasm -g -c output> <kernel-substr> <ncu source csv> <physics.cuh> [top]
Roles are separated by the BAR.SYNC layout: code is contiguous per role; the role of an address range is
recognised by which source functions dominate it (split_lookout_step / split_geometry_step / split_dynamics_step)."""
import csv, re, sys, collections
dis, fn, srccsv, cuh = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and fn in l and l.endswith(":"))
addr = {}; cur = None; chain = []
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith(".text."): break
    m = re.match(r'\s*//## File "(.*)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m: addr[int(m.group(1), 16)] = (cur, m.group(2))
src = open(cuh).read().split("\n")
rows = list(csv.reader(open(srccsv))); hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
stall_keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = None; recs = []
for r in rows[2:]:
    a = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
    if base is None: base = a
    key, ins = addr.get(a - base, (None, ""))
    recs.append({"off": a - base, "key": key, "ins": ins, "ex": int(r[ci["Instructions Executed"]] or 0),
                 "smp": int(r[ci["# Samples"]] or 0), "stall": {k: int(r[ci[k]] or 0) for k in stall_keys}})
# segment by barriers / EXIT: role = by counting executed per instruction (helpers run nsteps times too) - use heuristics:
# lookout code precedes geometry precedes dynamics in this build; boundaries = EXIT instructions
exits = [i for i, r in enumerate(recs) if r["ins"].startswith("EXIT") or " EXIT" in r["ins"]]
print("EXIT at instruction index:", exits[:12])
if len(sys.argv) > 6:
    lo, hi = [int(x) for x in sys.argv[6].split(":")]
    sel = recs[lo:hi]
    tot_ex = sum(r["ex"] for r in sel); tot_s = sum(r["smp"] for r in sel)
    print(f"range {lo}:{hi}: executed {tot_ex}, samples {tot_s}")
    agg = collections.Counter()
    for r in sel:
        for k, v in r["stall"].items(): agg[k] += v
    print({k: round(v / max(1, tot_s), 3) for k, v in agg.most_common(8)})
    byline = collections.defaultdict(lambda: [0, 0])
    for r in sel:
        byline[r["key"]][0] += r["ex"]; byline[r["key"]][1] += r["smp"]
    print("top source lines by samples:")
    for k, (e, sm) in sorted(byline.items(), key=lambda kv: -kv[1][1])[:top]:
        text = src[k[1] - 1].strip()[:100] if k and k[0] == "so101_physics.cuh" else str(k)
        print(f"  smp {100*sm/tot_s:5.2f}%  exec {100*e/tot_ex:5.2f}%  L{k[1] if k else 0:5d}  {text}")
