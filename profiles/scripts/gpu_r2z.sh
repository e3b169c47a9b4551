#!/bin/bash
# r2z: dynamic hand-out of the elements inside a launch against the grid-stride distribution
O=gpurun_out/r2z; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -q -x -k "fixture or reproducible or full_size or tiny or ghost" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for st in 0 1 0 1; do
GOMA_GPU_STATIC=$st python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_static${st}_$RANDOM.json 2>> $O/bench.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2z/bench_static*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); c=d["configs"]
            print(f.split("/")[-1], round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), {k: round(v["ms_per_step"],2) for k,v in c.items()})
PY
