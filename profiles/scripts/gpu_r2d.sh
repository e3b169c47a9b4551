#!/bin/bash
# round 2, fourth GPU call: 3x3 tensor-core blocks + scalar remainder tiles; attribution runs; chunked sweeps for hex8
O=gpurun_out/r2d; mkdir -p $O
( timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log )
tail -4 $O/pytest.log
B="--no-e2e --no-cpu-baseline --no-extra-configs --steps 3 --warmup 3"
timeout 600 python bench.py $B > $O/bench_c2.json 2> $O/bench_c2.err; python -c "import json;d=json.load(open('$O/bench_c2.json'));print('c2',d['ms_per_step'],d['roofline']['frac'])"
for dbg in 1 2 3; do GOMA_GPU_DEBUG=$dbg timeout 600 python bench.py $B > $O/bench_c2_debug$dbg.json 2>/dev/null; python -c "import json;d=json.load(open('$O/bench_c2_debug$dbg.json'));print('c2 debug $dbg',d['ms_per_step'])"; done
timeout 900 python bench.py --config c3 $B > $O/bench_c3.json 2> $O/bench_c3.err; python -c "import json;d=json.load(open('$O/bench_c3.json'));print('c3',d['ms_per_step'],d['roofline']['frac'])"
for ch in -1 0 40000 20000 10000 5000; do GOMA_GPU_CHUNK_ELEMS=$ch timeout 600 python bench.py --config c5 $B > $O/bench_c5_chunk$ch.json 2>/dev/null; python -c "import json;d=json.load(open('$O/bench_c5_chunk$ch.json'));print('c5 chunk $ch',d['ms_per_step'],d['gpu_launches'],d['roofline']['frac'])"; done
for dbg in 1 2; do GOMA_GPU_DEBUG=$dbg GOMA_GPU_CHUNK_ELEMS=-1 timeout 600 python bench.py --config c5 $B > $O/bench_c5_debug$dbg.json 2>/dev/null; python -c "import json;d=json.load(open('$O/bench_c5_debug$dbg.json'));print('c5 debug $dbg',d['ms_per_step'])"; done
for cfg in c2 c3; do
GOMA_GPU_LIB=$PWD/goma_b200/libgoma_gpu_fill_prof.so GOMA_GPU_PROFILE=1 timeout 300 python bench.py --config $cfg --edge 48 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/prof_$cfg.json 2> $O/prof_$cfg.err; grep "goma_gpu profile" $O/prof_$cfg.err | tail -2
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s 12 -c 1 -f -o $O/fill_c2 python bench.py --edge 64 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_c2.log 2>&1; tail -1 $O/ncu_c2.log
ls $O
