#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out/r2_sweep26.log
: > $out
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
python - >> $out <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_bench_default.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'steps', d['steps'], 'warmup', d['warmup'], 'fly', d['config']['batches_in_flight'], 'e2e', d['e2e']['value'], 'serial', d['one_batch_at_a_time']['value'])
print('roofline', d['roofline']['kernel'][:60], d['roofline']['frac'], d['roofline'].get('in_flight', {}).get('frac'))
print('encoder', d['encoder']['config2_ragged']['value'], d['encoder']['uniform']['value'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
PY
tail -4 gpurun_out/r2_bench_default.err >> $out
cat $out
