#!/bin/bash
# r2s: row-sum scaling with one exposed DRAM latency per batch; rows per CTA 4 / 8 / 2
mkdir -p gpurun_out/r2s
python -m pytest tests -q -m gpu -x -k "row_sum or csr or scale or post" > gpurun_out/r2s/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s/pytest.log
tail -2 gpurun_out/r2s/pytest.log
for r in 4 8 2; do
  GOMA_GPU_RSS_ROWS=$r python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2s/bench_rows$r.json 2> gpurun_out/r2s/bench_rows$r.err
python - <<PY
import json
for l in open("gpurun_out/r2s/bench_rows$r.json"):
    if l.startswith("{"):
        d=json.loads(l); p=d.get("post_fill"); c=d["configs"].get("c3_csr_layout")
        print($r, d["ms_per_step"], p["row_sum_scale_ms"], p["hbm_frac"], c["row_sum_scale_ms"], c["row_sum_scale_hbm_frac"], p["csr_values_ms"])
PY
done
