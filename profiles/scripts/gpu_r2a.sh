#!/bin/bash
# round 2, first GPU call: whole GPU suite, C2 / C3 bench lines, phase cycles, one full ncu capture of the C2 fill
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
( timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log )
tail -5 $O/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 600 $O/bench_c2.json
timeout 900 python bench.py --config c3 --edge 126 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 400 $O/bench_c3.json
GOMA_GPU_LIB=$PWD/goma_b200/libgoma_gpu_fill_prof.so GOMA_GPU_PROFILE=1 timeout 300 python bench.py --edge 48 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/prof_c2.json 2> $O/prof_c2.err; grep "goma_gpu profile" $O/prof_c2.err | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s 12 -c 1 -f -o $O/fill_c2 python bench.py --edge 64 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_c2.log 2>&1; tail -2 $O/ncu_c2.log
ls -la $O
