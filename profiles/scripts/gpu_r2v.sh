#!/bin/bash
# r2v: 2 GPUs, final round-2 library: the exchange tests on two devices, the default N = 2 line (weak, as the driver
# launches it, e2e included) and the strong-scaling line
O=gpurun_out/r2v; mkdir -p $O
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@"; }
timeout 600 python -m pytest tests/test_gpu_exchange.py -q -m gpu > $O/pytest_exchange.log 2>&1; echo "exchange tests rc=$?"; tail -2 $O/pytest_exchange.log
timeout 900 bash -c "$(declare -f run); run 2 --steps 3 --warmup 3" > $O/weak_c2_n2.json 2> $O/weak_c2_n2.err
timeout 900 bash -c "$(declare -f run); run 2 --scaling strong --partition brick --no-e2e --no-cpu-baseline --steps 5 --warmup 3" > $O/strong_c2_brick_n2.json 2> $O/strong_c2_brick_n2.err
timeout 300 bash -c "$(declare -f run); run 2 --impl reference --steps 2 --warmup 1" > $O/reference_n2.json 2> $O/reference_n2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2v/*.json')):
    d = None
    for line in open(f):
        if line.startswith('{'):
            d = json.loads(line)
    if d:
        print(f.split('/')[-1], d.get('n_gpus'), d.get('scaling'), round(d['value'] / 1e6, 3), 'M el/s', d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))
    else:
        print(f, 'NO JSON')
PY
tail -n 2 $O/*.err | grep -v "^$" | tail -12
