#!/bin/bash
# round 2: the final kernels -- GPU suite, the default bench line (headline + configs), captures and launch lists
O=gpurun_out/r2j; mkdir -p $O
( timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log )
tail -3 $O/pytest.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 200 $O/bench_default.json; tail -3 $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 300 $O/bench_reference.json
for cfg in c2 c3 c5 c4; do
e=64; [ $cfg = c3 ] && e=48; [ $cfg = c4 ] && e=24
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s 12 -c 1 -f -o $O/fill_$cfg python bench.py --config $cfg --edge $e --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_$cfg.log 2>&1; tail -1 $O/ncu_$cfg.log
done
timeout 600 ncu --set full --clock-control none -k regex:row_sum_scale -c 1 -f -o $O/row_sum_scale python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_rss.log 2>&1; tail -1 $O/ncu_rss.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/launches_default.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
ls $O
