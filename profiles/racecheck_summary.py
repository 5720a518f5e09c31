"""Histogram of a compute-sanitizer racecheck log: hazards by (kind, writer location, reader location).
usage: python profiles/racecheck_summary.py gpurun_out/sanitizer_racecheck.log"""
import collections
import re
import sys

kind, w, rows = None, None, collections.Counter()
for line in open(sys.argv[1], errors="replace"):
    m = re.search(r"(Error|Warning): (.*?hazard detected.*?) at (__shared__|__global__)", line)
    if m:
        kind = m.group(2).strip()
        continue
    m = re.search(r"(Write|Read) Thread \([0-9,]+\) at (.*)", line)
    if m and kind:
        loc = re.sub(r"\+0x[0-9a-f]+", "", m.group(2)).strip()
        loc = re.sub(r"\(CUtensorMap_st.*?\)", "(..)", loc)
        if w is None:
            w = (m.group(1), loc)
        else:
            rows[(kind, w, (m.group(1), loc))] += 1
            kind, w = None, None
for (k, a, b), n in rows.most_common():
    print(f"{n:8d}  {k}\n          {a[0]:5s} {a[1][:150]}\n          {b[0]:5s} {b[1][:150]}")
for line in open(sys.argv[1], errors="replace"):
    if "SUMMARY" in line:
        print(line.strip())
