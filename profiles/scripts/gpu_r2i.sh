#!/bin/bash
# round 2: vector stores in every first-touch write-out; default = 16 tensor-core blocks, two CTAs per SM
O=gpurun_out/r2i; mkdir -p $O
( timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log )
tail -3 $O/pytest.log
B="--no-e2e --no-cpu-baseline --no-extra-configs --steps 3 --warmup 3"
for v in 0 100 108 3; do
( GOMA_GPU_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fixture and (c2_hex27 or c3_hex27 or irr_hex27) or csr_layout and (c2_hex27 or c3_hex27)" > $O/pytest_var$v.log 2>&1; echo "variant $v parity rc=$?"; tail -1 $O/pytest_var$v.log )
GOMA_GPU_VARIANT=$v timeout 600 python bench.py $B > $O/bench_c2_var$v.json 2>/dev/null; python -c "import json;d=json.load(open('$O/bench_c2_var$v.json'));print('c2 variant $v',round(d['ms_per_step'],2),round(d['roofline']['frac'],3))"
GOMA_GPU_VARIANT=$v timeout 900 python bench.py --config c3 $B > $O/bench_c3_var$v.json 2>/dev/null; python -c "import json;d=json.load(open('$O/bench_c3_var$v.json'));print('c3 variant $v',round(d['ms_per_step'],2),round(d['roofline']['frac'],3))"
done
