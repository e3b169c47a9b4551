#!/bin/bash
# r2p: hex8 (C5) with 1 / 2 / 4 elements per CTA (Cfg::EPC)
mkdir -p gpurun_out/r2p
for e in 2 4; do
GOMA_GPU_EPC=$e python -m pytest tests -q -m gpu -x -k "c5 or hex8 or HEX8" > gpurun_out/r2p/pytest_$e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p/pytest_$e.log
tail -2 gpurun_out/r2p/pytest_$e.log
done
for e in 1 2 4; do
  GOMA_GPU_EPC=$e python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs > gpurun_out/r2p/bench_c5_epc$e.json 2> gpurun_out/r2p/bench_c5_epc$e.err
python - <<PY
import json
for l in open("gpurun_out/r2p/bench_c5_epc$e.json"):
    if l.startswith("{"):
        d=json.loads(l); print($e, d["ms_per_step"], d["device_ms_per_step"], d["roofline"]["frac"], d["gpu_launches"])
PY
tail -1 gpurun_out/r2p/bench_c5_epc$e.err
done
