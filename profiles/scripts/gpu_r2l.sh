#!/bin/bash
# r2l: row-sum scaling kernel, 16-byte staged copies, rows per CTA A/B
mkdir -p gpurun_out/r2l
python -m pytest tests -q -m gpu -x -k "row_sum or csr or scale or post" > gpurun_out/r2l/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l/pytest.log
for r in 8 4 2; do
  GOMA_GPU_RSS_ROWS=$r python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2l/bench_rows$r.json 2> gpurun_out/r2l/bench_rows$r.err
done
tail -3 gpurun_out/r2l/pytest.log
for r in 8 4 2; do python - <<PY
import json
for l in open("gpurun_out/r2l/bench_rows$r.json"):
    if l.startswith("{"):
        d=json.loads(l); print($r, d["ms_per_step"], d.get("post_fill"), d["configs"].get("c3_csr_layout"))
PY
done
