"""profiles/ncu_traffic.json from an `ncu --page raw --csv` export of one step of the default workload:

    python scripts/ncu_traffic.py gpurun_out/r2_26_prof_raw.csv "profiles/r2_26_ncu_all_kernels.txt" [key]

DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the two pool kernels; bench.py quotes them as
`roofline.traffic` for the workload whose key matches (B, M, D, H, dtype, dropout[, fold])."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    raw, source = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else "B=65536,M=3,D=512,H=8,bf16,dropout=0.0,fold"
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    found = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        which = "pool_bwd" if "pool_bwd_kernel" in name else ("pool_fwd" if "pool_fwd" in name else None)
        if which is None or which in found:
            continue
        total = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(r[idx[m]]) * UNIT[units[idx[m]]]
        found[which] = int(total)
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    found["_source"] = f"{source} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)"
    table[key] = found
    json.dump(table, open(path, "w"), indent=1)
    print(key, found)


if __name__ == "__main__":
    main()
