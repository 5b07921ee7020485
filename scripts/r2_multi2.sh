#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-2}
out=gpurun_out/r2_multi_n$N.log
: > $out
nvidia-smi -L >> $out
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3 >> $out; fi
echo "== weak scaling, N=$N, verify" >> $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --verify --no-extras --no-cpu-baseline 2>> gpurun_out/r2_multi_n$N.err | grep '^{' > gpurun_out/r2_weak_n$N.json
python - $N >> $out <<'PY'
import json, sys
n = sys.argv[1]
for name in (f"gpurun_out/r2_weak_n{n}.json",):
    try:
        d = json.loads(open(name).read().strip().split("\n")[-1])
        print(name, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "verify", d["verify"], "fly", d["config"]["batches_in_flight"])
    except Exception as e:
        print(name, "failed", e)
PY
if [ "$N" = "8" ]; then
  echo "== strong scaling: BASELINE configs[4], 65,536 equations" >> $out
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --total 65536 --verify --no-extras --no-cpu-baseline 2>> gpurun_out/r2_multi_n$N.err | grep '^{' > gpurun_out/r2_strong_n$N.json
  python - >> $out <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_strong_n8.json").read().strip().split("\n")[-1])
    print("strong", "value", d["value"], "ms", d["ms_per_step"], "steps", d["steps"], "e2e", d["e2e"]["value"], "verify", d["verify"], "scaling", d["scaling"])
except Exception as e:
    print("strong failed", e)
PY
fi
tail -c 600 gpurun_out/r2_multi_n$N.err >> $out
cat $out
