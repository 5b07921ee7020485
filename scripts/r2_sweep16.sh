#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep16.log
: > $out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 >> $out
for i in 1 2; do
  echo "== bench K=20 W=5 run $i" >> $out
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>> $out | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'serial', (d.get('one_batch_at_a_time') or {}).get('value'), 'fly', d['config']['batches_in_flight'], d['clocks'])
" >> $out
done
B="decode_branches=1"
for opt in "$B" "$B,gemm_min_ctas=60" ; do
  timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$opt" >> $out 2>&1 || echo "FAILED $opt" >> $out
done
cat $out
