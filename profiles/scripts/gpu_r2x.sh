#!/bin/bash
# r2x: w = A v from the node-level neighbour lists (goma_gpu_matvec): parity, then the timing in the post-fill leg
O=gpurun_out/r2x; mkdir -p $O
python -m pytest tests -q -m gpu -x -k "matvec" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs > $O/bench.json 2> $O/bench.err
python - <<PY
import json
for l in open("gpurun_out/r2x/bench.json"):
    if l.startswith("{"):
        d=json.loads(l); p=d.get("post_fill"); print(d["ms_per_step"], {k:v for k,v in p.items() if "matvec" in k or "row_sum" in k})
PY
tail -2 $O/bench.err
