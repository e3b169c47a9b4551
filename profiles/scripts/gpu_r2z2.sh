#!/bin/bash
# r2z2: share of a launch handed out by block index before the counter takes over
O=gpurun_out/r2z2; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -q -x -k "fixture and first_touch or reproducible or tiny" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest.log
for pct in 60 75 85 90; do  # (run as two calls: 75 90, then 85 60)
GOMA_GPU_STATIC_PCT=$pct python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_pct$pct.json 2>> $O/bench.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2z2/bench_pct*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); c=d["configs"]
            print(f.split("/")[-1], round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), {k: round(v["ms_per_step"],2) for k,v in c.items()})
PY
