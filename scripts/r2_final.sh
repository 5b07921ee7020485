#!/bin/bash
# round-2 closing run: full GPU suite, smoke(), the bench line as the driver runs it, the reference arm, refreshed ncu launch list
cd "$(dirname "$0")/.."
out=gpurun_out/r2_final.log
: > $out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 >> $out
echo "== smoke" >> $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 >> $out
echo "== bench (driver flags)" >> $out
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; tail -c 400 gpurun_out/r2_bench3.err >> $out
python - >> $out <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2_bench3.json') if l.startswith('{')][-1])
    for k in ('value','ms_per_step','e2e','one_batch_at_a_time','gpu_launches','clocks','roofline','roofline_attention','job_hbm','kernel_time_shares','cpu_baseline','encoder'):
        print(k, json.dumps(d.get(k))[:1400])
    print('by_class', json.dumps(d.get('roofline_by_class'))[:3000])
except Exception as e:
    print('bench parse failed', e)
PY
echo "== reference arm" >> $out
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -2 >> $out
echo "== ncu launch list" >> $out
CMD2="python scripts/profile_generate.py --batch 512 --max-len 48 --warm 0 --no-graph --branches 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_v3.csv $CMD2 > gpurun_out/r2_launches_v3.log 2>&1
ls -la gpurun_out/r2_launches_v3.csv >> $out
cat $out
