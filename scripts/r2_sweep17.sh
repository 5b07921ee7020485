#!/bin/bash
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
out=gpurun_out/r2_sweep17.log
: > $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -x -q -m gpu -k "encoder or groupnorm or im2col or bf16_tier or smoke" 2>&1 | tail -6 >> $out
timeout 200 python scripts/encoder_profile.py 512 >> $out 2>&1
echo "== bench K=20 W=5" >> $out
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>> $out | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'serial', (d.get('one_batch_at_a_time') or {}).get('value'), 'fly', d['config']['batches_in_flight'], d['clocks'])
" >> $out
cat $out
