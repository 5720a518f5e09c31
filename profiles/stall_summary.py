"""Summarise an `ncu --page source --csv` dump: top SASS lines by stall samples and totals per stall reason.
usage: ncu -i X.ncu-rep --page source --csv --kernel-id ::regex:<name>:<n> > src.csv ; python profiles/stall_summary.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hdr_i], [r for r in rows[hdr_i + 1:] if len(r) >= len(rows[hdr_i]) - 2]
print(rows[0][:2])
ia, isamp = hdr.index("Source"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]


def val(r, i):
    try:
        return int(r[i])
    except Exception:
        return 0


tot = sum(val(r, isamp) for r in data)
print("total samples", tot)
for r in sorted(data, key=lambda r: -val(r, isamp))[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    st = sorted([(val(r, i), hdr[i]) for i in stall_cols], reverse=True)[:2]
    print(f"{val(r, isamp):6d} {100.0 * val(r, isamp) / max(tot, 1):5.1f}%  {r[ia].strip()[:72]:72s} {st}")
agg = {hdr[i]: sum(val(r, i) for r in data) for i in stall_cols}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
