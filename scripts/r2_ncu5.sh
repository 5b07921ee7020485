#!/bin/bash
cd "$(dirname "$0")/.."
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 600 $NCU -k 'regex:tc_stem_kernel' -c 1 -o gpurun_out/r2b_stem -f python scripts/encoder_only.py 512 1 > gpurun_out/r2b_ncu_stem.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2b_stem.ncu-rep > gpurun_out/r2b_stem_summary.md 2>&1
ncu -i gpurun_out/r2b_stem.ncu-rep --page details --csv 2>/dev/null | grep -i -E "stall|Issue|Eligible|No Eligible|One or More" | head -40 >> gpurun_out/r2b_stem_summary.md
ncu -i gpurun_out/r2b_stem.ncu-rep --page source --csv 2>/dev/null > gpurun_out/r2b_stem_source.csv
python - <<'PY' >> gpurun_out/r2b_stem_summary.md
import csv
rows = list(csv.reader(open("gpurun_out/r2b_stem_source.csv")))
hdr = rows[0]
print(hdr[:12])
# top source lines by sampled stalls
try:
    si = [i for i, h in enumerate(hdr) if "Samples" in h][0]
    src = [i for i, h in enumerate(hdr) if h in ("Source", "SASS")][0]
    top = sorted(rows[1:], key=lambda r: -float(r[si] or 0))[:25]
    for r in top: print(r[si], r[src][:150])
except Exception as e:
    print("source parse failed", e)
PY
rm -f gpurun_out/r2b_stem.ncu-rep gpurun_out/r2b_stem_source.csv
cat gpurun_out/r2b_stem_summary.md | cut -c1-300
