#!/bin/bash
# r2n: row-sum scaling, batches cut by entry count (E entries + one row of slack) against 4 rows per CTA
mkdir -p gpurun_out/r2n
for e in 2048 1024; do
GOMA_GPU_RSS_ENTRIES=$e python -m pytest tests -q -m gpu -x -k "row_sum or csr or scale or post" > gpurun_out/r2n/pytest_$e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n/pytest_$e.log
tail -2 gpurun_out/r2n/pytest_$e.log
done
for e in 0 1024 2048 4096; do
  GOMA_GPU_RSS_ENTRIES=$e python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2n/bench_e$e.json 2> gpurun_out/r2n/bench_e$e.err
python - <<PY
import json
for l in open("gpurun_out/r2n/bench_e$e.json"):
    if l.startswith("{"):
        d=json.loads(l); p=d.get("post_fill"); c=d["configs"].get("c3_csr_layout")
        print($e, d["ms_per_step"], p["row_sum_scale_ms"], p["hbm_frac"], c["row_sum_scale_ms"], c["row_sum_scale_hbm_frac"])
PY
done
