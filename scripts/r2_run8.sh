#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out/r2_run8.log
: > $out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 >> $out
echo "== bench" >> $out
timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; tail -c 600 gpurun_out/r2_bench1.err >> $out
python - >> $out <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2_bench1.json') if l.startswith('{')][-1])
    for k in ('value','ms_per_step','e2e','one_batch_at_a_time','gpu_launches','clocks','roofline','job_hbm','kernel_time_shares','cpu_baseline','encoder'):
        print(k, json.dumps(d.get(k))[:900])
    print('by_class', json.dumps(d.get('roofline_by_class'))[:3000])
except Exception as e:
    print('bench parse failed', e)
PY
echo "== reference arm" >> $out
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 >> $out 2>&1
cat $out
