"""Opcode histogram (executed warp instructions) of an `ncu --page source --csv` export:
python scripts/sass_histogram.py src.csv [units]   (units = what to normalise by, e.g. 65536 samples)"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
i_src, i_ex, i_s = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp, total = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    try:
        n, s = int(r[i_ex]), int(r[i_s])
    except (ValueError, IndexError):
        continue
    tok = r[i_src].split()
    if not tok:
        continue
    o = tok[1] if tok[0].startswith("@") else tok[0]
    parts = o.rstrip(";").split(".")
    o = parts[0] + ("." + parts[1] if parts[0] in ("MUFU", "F2F", "I2F", "F2I", "F2FP", "LDS", "STS", "LDG", "STG", "LDGSTS") and len(parts) > 1 else "")
    ops[o] += n
    samp[o] += s
    total += n
print(rows[0][1][:100])
print(f"total warp instructions {total}  ({total / units:.1f} per unit)")
for o, n in ops.most_common(30):
    print(f"  {o:18s} {n:11d} {n / units:8.2f}/unit   stall samples {samp[o]}")
