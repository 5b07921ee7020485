#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out/r2_run10.log
: > $out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 >> $out
echo "== bench" >> $out
timeout 1200 python bench.py --steps 12 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -c 400 gpurun_out/r2_bench2.err >> $out
python - >> $out <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2_bench2.json') if l.startswith('{')][-1])
    for k in ('value','ms_per_step','e2e','one_batch_at_a_time','gpu_launches','clocks','roofline','roofline_attention','job_hbm','kernel_time_shares','cpu_baseline','encoder'):
        print(k, json.dumps(d.get(k))[:1200])
except Exception as e:
    print('bench parse failed', e)
PY
cat $out
