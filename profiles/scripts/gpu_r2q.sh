#!/bin/bash
# r2q: cache operator of the first-touch stores: default (write-back) | .cs (evict-first) | .wt (write-through)
mkdir -p gpurun_out/r2q
R=$PWD
for v in "" _cs _wt; do
  export GOMA_GPU_LIB=$R/goma_b200/libgoma_gpu_fill$v.so
  python -m pytest tests/test_gpu_parity.py -q -x -k "fixture and first_touch" > gpurun_out/r2q/pytest$v.log 2>&1; echo "lib$v parity rc=$?"
  for c in c2 c5; do
    python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs > gpurun_out/r2q/bench_$c$v.json 2> gpurun_out/r2q/bench_$c$v.err
    python - <<PY
import json
for l in open("gpurun_out/r2q/bench_$c$v.json"):
    if l.startswith("{"):
        d=json.loads(l); print("lib$v", "$c", round(d["ms_per_step"],2), round(d["roofline"]["frac"],3))
PY
  done
done
