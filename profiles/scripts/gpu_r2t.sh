#!/bin/bash
# r2t: csr_values with the next row's descriptors prefetched
mkdir -p gpurun_out/r2t
python -m pytest tests -q -m gpu -x -k "csr or handoff" > gpurun_out/r2t/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t/pytest.log
tail -2 gpurun_out/r2t/pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs > gpurun_out/r2t/bench.json 2> gpurun_out/r2t/bench.err
python - <<PY
import json
for l in open("gpurun_out/r2t/bench.json"):
    if l.startswith("{"):
        d=json.loads(l); p=d.get("post_fill"); print(d["ms_per_step"], p)
PY
