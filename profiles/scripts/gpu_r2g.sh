#!/bin/bash
# round 2: A/B of the tensor-core kernel variants (9 blocks + scalar remainder | 16 padded blocks) x (3 CTAs at 80 regs | 2 at 128)
O=gpurun_out/r2g; mkdir -p $O
B="--no-e2e --no-cpu-baseline --no-extra-configs --steps 3 --warmup 3"
for v in 0 1 2 3; do
GOMA_GPU_VARIANT=$v timeout 600 python bench.py $B > $O/bench_c2_var$v.json 2>/dev/null; python -c "import json;d=json.load(open('$O/bench_c2_var$v.json'));print('c2 variant $v',round(d['ms_per_step'],2),round(d['roofline']['frac'],3))"
GOMA_GPU_VARIANT=$v timeout 900 python bench.py --config c3 $B > $O/bench_c3_var$v.json 2>/dev/null; python -c "import json;d=json.load(open('$O/bench_c3_var$v.json'));print('c3 variant $v',round(d['ms_per_step'],2),round(d['roofline']['frac'],3))"
done
