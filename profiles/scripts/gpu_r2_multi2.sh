#!/bin/bash
# round 2, 8 GPUs of one box: strong scaling of configs[1] with 2 / 2x2 / 2x2x2 BRICKS (7 neighbours per rank at N = 8),
# weak scaling of C5 (hex8, 100^3 per GPU -> 8M elements) without chunked sweeps
O=gpurun_out/r2multi2; mkdir -p $O
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@"; }
Q="--no-e2e --no-cpu-baseline --no-extra-configs --steps 5 --warmup 3"
timeout 600 python bench.py $Q > $O/strong_c2_n1.json 2> $O/strong_c2_n1.err
for n in 2 4 8; do timeout 900 bash -c "$(declare -f run); run $n --scaling strong --partition brick $Q" > $O/strong_c2_brick_n$n.json 2> $O/strong_c2_brick_n$n.err; done
timeout 600 python bench.py --config c5 $Q > $O/weak_c5_n1.json 2> $O/weak_c5_n1.err
timeout 900 bash -c "$(declare -f run); run 8 --config c5 $Q" > $O/weak_c5_n8.json 2> $O/weak_c5_n8.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2multi2/*.json')):
    d = None
    for line in open(f):
        if line.startswith('{'):
            d = json.loads(line)
    if d:
        print(f.split('/')[-1], d['n_gpus'], d['scaling'], round(d['value'] / 1e6, 2), 'M el/s', round(d['ms_per_step'], 2), 'ms', d['config'].get('neighbors'), d['config']['elements_assembled_per_gpu'])
    else:
        print(f, 'NO JSON')
PY
tail -n 3 $O/*.err | grep -v "^$" | tail -20
