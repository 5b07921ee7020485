#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep15.log
: > $out
for n in 6 7 10 5; do
  echo "== in-flight $n" >> $out
  timeout 300 python bench.py --steps 20 --warmup 5 --in-flight $n --no-cpu-baseline --no-extras 2>> $out | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'serial', (d.get('one_batch_at_a_time') or {}).get('value'))
" >> $out
done
B="decode_branches=1"
for opt in "$B" "$B,dbg_skip=4" "$B,pdl_mid=0" ; do
  timeout 300 python scripts/inflight_probe.py 512 256 6 6 "$opt" >> $out 2>&1 || echo "FAILED $opt" >> $out
done
cat $out
