#!/bin/bash
# round 2, third GPU call: greedy-order colouring on the device, CSR layout, batched row-sum scaling; captures of C5 / C4
O=gpurun_out/r2c; mkdir -p $O
( timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log )
tail -5 $O/pytest.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json; tail -3 $O/bench_default.err
for cfg in c2 c3; do
GOMA_GPU_LIB=$PWD/goma_b200/libgoma_gpu_fill_prof.so GOMA_GPU_PROFILE=1 timeout 300 python bench.py --config $cfg --edge 48 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/prof_$cfg.json 2> $O/prof_$cfg.err; grep "goma_gpu profile" $O/prof_$cfg.err | tail -2
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s 12 -c 1 -f -o $O/fill_c2 python bench.py --edge 64 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_c2.log 2>&1; tail -2 $O/ncu_c2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s 12 -c 1 -f -o $O/fill_c3 python bench.py --config c3 --edge 48 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_c3.log 2>&1; tail -2 $O/ncu_c3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s 12 -c 1 -f -o $O/fill_c5 python bench.py --config c5 --edge 64 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_c5.log 2>&1; tail -2 $O/ncu_c5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s 12 -c 1 -f -o $O/fill_c4 python bench.py --config c4 --edge 24 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_c4.log 2>&1; tail -2 $O/ncu_c4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/launches_c2.log 2>&1
ls -la $O
