#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out/r2_sweep24.log
: > $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sampling.py -x -q -m gpu 2>&1 | tail -5 >> $out
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>> $out | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'serial', d.get('one_batch_at_a_time'))
" >> $out
cat $out
