#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel the instruction histogram (top mnemonics) and every line that shows a
Blackwell / Hopper-class feature (UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDG.E.*256 = 256-bit loads, REDUX, VIMNMX3).
Usage: python tools/sass_summary.py [kernel-substring ...] > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "splintr_b200", "libsplintr_b200.so")
want = sys.argv[1:] or ["k_probe", "k_bpe", "k_emit", "k_pretok_fast"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, hist, feat = None, collections.defaultdict(collections.Counter), collections.defaultdict(list)
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if cur and m:
        op = m.group(2)
        hist[cur][op.split(".")[0]] += 1
        if re.search(r"UBLKCP|SYNCS|\.256|REDUX|VIMNMX3|UTMA|ELECT", op):
            feat[cur].append(ln.split("/*")[1][:4] + " " + ln.split("*/")[1].split(";")[0].strip())
for k in sorted(hist):
    if not any(w in k for w in want):
        continue
    h = hist[k]
    print(f"== {k}: {sum(h.values())} instructions")
    print("   " + ", ".join(f"{op} {n}" for op, n in h.most_common(14)))
    seen = collections.Counter(re.sub(r"^@!?U?P\d+ ", "", x.split(" ", 1)[1]).split(" ")[0] for x in feat[k])
    print("   features: " + (", ".join(f"{op} x{n}" for op, n in seen.items()) or "-"))
    for x in feat[k][:12]:
        print("      " + x)
