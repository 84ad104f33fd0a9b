#!/usr/bin/env python
"""Per-source-line warp-stall breakdown of one kernel of an ncu report (joins the source page with nvdisasm line info).
usage: tools/ncu_stalls.py <report.ncu-rep> <object with the kernel> <kernel substring> [top N]"""
import csv, glob, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 12
obj = os.path.abspath(obj)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
lines_of = {}
for cb in glob.glob(tmp + "/*.cubin"):
    dis = subprocess.run(["nvdisasm", "-gi", cb], capture_output=True, text=True).stdout.splitlines()
    inside, pending, cur = False, [], [("?", 0)]
    for ln in dis:
        if ln.startswith(".text."):
            inside = kern in ln
            pending = []
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            pending.append((m.group(1).split("/")[-1], int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
        if m:
            if pending:
                cur, pending = list(pending), []
            lines_of[int(m.group(1), 16)] = (cur, m.group(2))
    if lines_of:
        break
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True,
                     text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ia, ii = hdr.index("Address"), hdr.index("Instructions Executed")
body = [r for r in rows[hi + 1:] if r and r[0].startswith("0x")]
base = int(body[0][ia], 16)
tot = defaultdict(int)
by_line = defaultdict(lambda: defaultdict(int))
inst_line = defaultdict(int)
for r in body:
    off = int(r[ia], 16) - base
    chain, _ = lines_of.get(off, ([("?", 0)], ""))
    key = chain[0]
    inst_line[key] += int(r[ii] or 0)
    for i, h in stall_cols:
        v = int(r[i] or 0)
        tot[h] += v
        by_line[h][key] += v
allsamp = sum(tot.values())
print(f"kernel {kern}: {allsamp} stall samples, {sum(inst_line.values())} warp-inst")
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v * 100 < allsamp:
        continue
    print(f"\n== {h}: {100.0 * v / allsamp:.1f} % of samples ==")
    for k, c in sorted(by_line[h].items(), key=lambda kv: -kv[1])[:top]:
        print(f"   {k[0] + ':' + str(k[1]):30s} {100.0 * c / allsamp:6.2f} %   (inst {100.0 * inst_line[k] / max(sum(inst_line.values()), 1):5.2f} %)")
