#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep27.log
: > $out
for n in 5 6 7 10 5 6; do
  timeout 300 python bench.py --steps 20 --warmup 5 --in-flight $n --no-cpu-baseline --no-extras 2>> $out | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('in-flight', d['config']['batches_in_flight'], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value']))
" >> $out
done
cat $out
