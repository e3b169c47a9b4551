#!/bin/bash
# round 2: one unified operand table (w folded into the B fragments), dphi from 1-D factors, three CTAs per SM
O=gpurun_out/r2f; mkdir -p $O
( timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log )
tail -4 $O/pytest.log
B="--no-e2e --no-cpu-baseline --no-extra-configs --steps 3 --warmup 3"
timeout 600 python bench.py $B > $O/bench_c2.json 2> $O/bench_c2.err; python -c "import json;d=json.load(open('$O/bench_c2.json'));print('c2',d['ms_per_step'],d['roofline']['frac'])"
for dbg in 1 2 3; do GOMA_GPU_DEBUG=$dbg timeout 600 python bench.py $B > $O/bench_c2_debug$dbg.json 2>/dev/null; python -c "import json;d=json.load(open('$O/bench_c2_debug$dbg.json'));print('c2 debug $dbg',d['ms_per_step'])"; done
for mb in 3 2; do GOMA_GPU_C3_MINB=$mb timeout 900 python bench.py --config c3 $B > $O/bench_c3_minb$mb.json 2> $O/bench_c3.err; python -c "import json;d=json.load(open('$O/bench_c3_minb$mb.json'));print('c3 minb $mb',d['ms_per_step'],d['roofline']['frac'])"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s 12 -c 1 -f -o $O/fill_c2 python bench.py --edge 64 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_c2.log 2>&1; tail -1 $O/ncu_c2.log
ls $O
