#!/usr/bin/env python
"""Joins an ncu report's per-SASS-instruction counters with nvdisasm line info and prints, per source line,
the executed warp instructions, thread instructions and stall samples of one kernel.

usage: tools/ncu_lines.py <report.ncu-rep> <object-or-so with the kernel> <kernel substring> [top N]
Lines are attributed twice: to the innermost location and to the outermost location inside `--file` (default
fb_fused.cuh), so both "which helper is hot" and "which call site is hot" can be read off."""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    import os
    rep, obj, kern = sys.argv[1:4]
    obj = os.path.abspath(obj)
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    focus = sys.argv[5] if len(sys.argv) > 5 else "fb_fused.cuh"
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
    import glob
    cubins = glob.glob(tmp + "/*.cubin")
    lines_of = {}
    for cb in cubins:
        dis = subprocess.run(["nvdisasm", "-gi", cb], capture_output=True, text=True).stdout.splitlines()
        inside = False
        pending = []
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
                off = int(m.group(1), 16)
                if pending:
                    cur = list(pending)
                    pending = []
                lines_of[off] = (cur, m.group(2))
        if lines_of:
            break
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True,
                         text=True).stdout.splitlines()
    rows = list(csv.reader(raw))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index(
        "Thread Instructions Executed"), hdr.index("# Samples")
    body = [r for r in rows[hi + 1:] if r and r[0].startswith("0x")]
    base = int(body[0][ia], 16)
    inner = defaultdict(lambda: [0, 0, 0])
    outer = defaultdict(lambda: [0, 0, 0])
    depth1 = defaultdict(lambda: [0, 0, 0])
    depth2 = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in body:
        off = int(r[ia], 16) - base
        chain, _ = lines_of.get(off, ([("?", 0)], ""))
        vals = [int(r[ii] or 0), int(r[it] or 0), int(r[isamp] or 0)]
        k_in = chain[0]
        foc = [c for c in chain if c[0] == focus]
        k_out = foc[-1] if foc else chain[-1]
        # one and two inlining levels below the outermost location
        io = max(i for i, c in enumerate(chain) if c == k_out)
        k_d1 = chain[max(io - 1, 0)]
        k_d2 = chain[max(io - 2, 0)]
        for d, k in ((inner, k_in), (outer, k_out), (depth1, k_d1), (depth2, k_d2)):
            for j in range(3):
                d[k][j] += vals[j]
        for j in range(3):
            tot[j] += vals[j]
    print(f"kernel {kern}: warp-inst {tot[0]}, thread-inst {tot[1]}, samples {tot[2]}")
    for name, d in (("innermost location", inner), (f"outermost location in {focus}", outer),
                    ("one inlining level below the outermost", depth1), ("two levels below", depth2)):
        print(f"\n== by {name} ==")
        print(f"{'file:line':32s} {'warp-inst':>12s} {'%':>6s} {'lanes':>6s} {'samples%':>8s}")
        for k, v in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"{k[0] + ':' + str(k[1]):32s} {v[0]:12d} {100.0 * v[0] / max(tot[0], 1):6.2f} "
                  f"{v[1] / max(v[0], 1):6.1f} {100.0 * v[2] / max(tot[2], 1):8.2f}")


if __name__ == "__main__":
    main()
