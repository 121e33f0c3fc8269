#!/usr/bin/env python
"""Condense a compute-sanitizer --tool racecheck log: hazards grouped by (kind, writer site, reader site) with counts, plus the
tool's own summary and the pytest verdict.  The raw log is tens of MB (one backtrace per hazard)."""
import re
import sys
from collections import Counter

lines = open(sys.argv[1], errors="replace").read().splitlines()
groups, cur = Counter(), None
for i, l in enumerate(lines):
    m = re.match(r"========= (Error|Warning): (.*?) at __shared__ 0x[0-9a-f]+ in block", l)
    if m:
        site = lambda s: re.sub(r"\+0x[0-9a-f]+", "", re.sub(r"^=========\s+(Write|Read) Thread \(\d+,\d+,\d+\)( \(block rank \d\))? at ", "", s)).strip()
        w = site(lines[i + 1]) if i + 1 < len(lines) else "?"
        r = site(lines[i + 2]) if i + 2 < len(lines) else "?"
        groups[(m.group(2), w, r)] += 1
print("# compute-sanitizer --tool racecheck --racecheck-report all: hazards by (kind, first access, second access)")
for (kind, w, r), n in groups.most_common():
    print(f"{n:6d}  {kind}\n          first : {w}\n          second: {r}")
for l in lines:
    if "RACECHECK SUMMARY" in l or re.search(r"\d+ (passed|failed)", l):
        print(l)
