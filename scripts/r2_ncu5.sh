#!/bin/bash
cd "$(dirname "$0")/.."
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 600 $NCU -k 'regex:tc_stem_kernel' -c 1 -o gpurun_out/r2b_stem -f python scripts/encoder_only.py 512 1 > gpurun_out/r2b_ncu_stem.log 2>&1
ncu -i gpurun_out/r2b_stem.ncu-rep --page source --csv 2>/dev/null > gpurun_out/r2b_stem_source.csv
python - <<'PY' > gpurun_out/r2b_stem_summary.md
import csv
rows = list(csv.reader(open("gpurun_out/r2b_stem_source.csv")))
print(len(rows), "rows")
hi = next((i for i, r in enumerate(rows) if any("Sampl" in c for c in r)), None)
print("header row", hi, rows[hi] if hi is not None else rows[:3])
if hi is not None:
    hdr = rows[hi]
    si = next(i for i, h in enumerate(hdr) if "Sampl" in h)
    def f(x):
        try: return float(x)
        except Exception: return 0.0
    body = [r for r in rows[hi + 1:] if len(r) > si]
    tot = sum(f(r[si]) for r in body)
    print("total samples", tot, "column", hdr[si])
    for r in sorted(body, key=lambda r: -f(r[si]))[:45]:
        print(r[si], "|", " | ".join(c[:120] for c in r[:3]))
PY
rm -f gpurun_out/r2b_stem.ncu-rep gpurun_out/r2b_stem_source.csv
cut -c1-260 gpurun_out/r2b_stem_summary.md
