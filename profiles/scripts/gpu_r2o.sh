#!/bin/bash
# r2o: host streaming (goma_gpu_problem.host_stream_chunks) -- parity test, then e2e with 0 / 4 / 8 / 16 chunks
mkdir -p gpurun_out/r2o
python -m pytest tests -q -m gpu -x -k "host_stream" > gpurun_out/r2o/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o/pytest.log
tail -5 gpurun_out/r2o/pytest.log
for k in 8 16 4; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-configs --host-stream-chunks $k > gpurun_out/r2o/bench_k$k.json 2> gpurun_out/r2o/bench_k$k.err
python - <<PY
import json
for l in open("gpurun_out/r2o/bench_k$k.json"):
    if l.startswith("{"):
        d=json.loads(l); e=d["e2e"]
        print($k, d["ms_per_step"], e["value"], e["ms_per_step"], e["streaming"], e["host_copy_roof"]["d2h_matrix_ms"])
PY
tail -2 gpurun_out/r2o/bench_k$k.err
done
