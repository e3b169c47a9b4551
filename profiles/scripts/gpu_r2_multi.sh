#!/bin/bash
# round 2, 8 GPUs of one box: strong scaling of configs[1] (ONE 100^3 hex27 mesh split over N GPUs), weak scaling of C5
# (hex8, 100^3 per GPU -> 8M elements), weak C2 with the host-buffer end-to-end leg (NUMA-bound ranks)
O=gpurun_out/r2multi; mkdir -p $O
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@"; }
Q="--no-e2e --no-cpu-baseline --no-extra-configs --steps 5 --warmup 3"
timeout 600 python bench.py $Q > $O/strong_c2_n1.json 2> $O/strong_c2_n1.err
for n in 2 4 8; do timeout 900 bash -c "$(declare -f run); run $n --scaling strong $Q" > $O/strong_c2_n$n.json 2> $O/strong_c2_n$n.err; done
for n in 1 2 4 8; do python -c "import json;d=json.load(open('$O/strong_c2_n$n.json'));print('strong c2 n=$n',round(d['value']/1e6,2),'M el/s',round(d['ms_per_step'],2),'ms')"; done
timeout 600 python bench.py --config c5 $Q > $O/weak_c5_n1.json 2> $O/weak_c5_n1.err
timeout 900 bash -c "$(declare -f run); run 8 --config c5 $Q" > $O/weak_c5_n8.json 2> $O/weak_c5_n8.err
for n in 1 8; do python -c "import json;d=json.load(open('$O/weak_c5_n$n.json'));print('weak c5 n=$n',round(d['value']/1e6,2),'M el/s',round(d['ms_per_step'],2),'ms')"; done
timeout 900 bash -c "$(declare -f run); run 8 --no-cpu-baseline --no-extra-configs --steps 5 --warmup 3 --e2e-steps 1" > $O/weak_c2_n8.json 2> $O/weak_c2_n8.err
python -c "import json;d=json.load(open('$O/weak_c2_n8.json'));print('weak c2 n=8',round(d['value']/1e6,2),'M el/s',round(d['ms_per_step'],2),'ms; e2e',round(d['e2e']['value']/1e6,2),'M el/s', d['e2e'].get('host_copy_roof',{}).get('GB/s'), d['config'].get('numa_binding'))"
tail -2 $O/*.err | tail -20
