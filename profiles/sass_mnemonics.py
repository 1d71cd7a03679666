#!/usr/bin/env python
"""Counts the Blackwell mnemonics in the shipped library, whole library and per kernel:  python profiles/sass_mnemonics.py > profiles/rNN_sass_mnemonics.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "neuralode4j_b200/libnode4j_b200.so"
WANT = ["ATOM", "LDGSTS", "LDTM", "STTM", "SYNCS", "UBLKCP", "UTMALDG", "UTCBAR", "UTCHMMA", "HMMA"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur[:cur.index("(")] if "(" in cur else cur
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        if op in WANT:
            per[cur][op] += 1
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print(f"# cuobjdump -sass {LIB}: Blackwell mnemonics, whole library and per kernel that carries tensor-core / bulk-copy instructions")
print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS.* = mbarrier, LDGSTS = cp.async")
print("total: " + ", ".join(f"{k} {v}" for k, v in sorted(tot.items())))
for k, c in per.items():
    if c.get("UTCHMMA") or c.get("UBLKCP") or c.get("LDTM"):
        print(f"{k}: " + ", ".join(f"{a} {b}" for a, b in sorted(c.items())))
