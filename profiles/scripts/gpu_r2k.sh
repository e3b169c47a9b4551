#!/bin/bash
# round 2: double-buffered row-sum scaling, exchange fence (GPU suite), launch list of the headline step
O=gpurun_out/r2k; mkdir -p $O
( timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log )
tail -3 $O/pytest.log
timeout 900 python bench.py --no-cpu-baseline --no-extra-configs --e2e-steps 1 > $O/bench_c2.json 2> $O/bench_c2.err; python -c "import json;d=json.load(open('$O/bench_c2.json'));print('c2',d['ms_per_step'],d['post_fill'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 760 -c 400 --csv --log-file $O/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/launches_c2.log 2>&1
grep -c fill_kernel $O/launches_c2.csv
timeout 600 ncu --set full --clock-control none -k regex:row_sum_scale -c 1 -f -o $O/row_sum_scale python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_rss.log 2>&1
ncu -i $O/row_sum_scale.ncu-rep --page raw --csv > $O/r2k_row_sum_scale_raw.csv 2>/dev/null; rm -f $O/row_sum_scale.ncu-rep
