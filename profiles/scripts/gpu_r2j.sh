#!/bin/bash
# round 2: the final kernels -- GPU suite, the default bench line (headline + configs), captures (summarised on the box:
# gpurun_out/ is capped at 64 MiB) and the launch list of the headline step
O=gpurun_out/r2j; mkdir -p $O
( timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log )
tail -3 $O/pytest.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 200 $O/bench_default.json; tail -3 $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 300 $O/bench_reference.json
for cfg in c2 c3 c5 c4; do
e=64; [ $cfg = c3 ] && e=48; [ $cfg = c4 ] && e=24
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s 12 -c 1 -f -o $O/fill_$cfg python bench.py --config $cfg --edge $e --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_$cfg.log 2>&1; tail -1 $O/ncu_$cfg.log
python profiles/summarize_ncu.py $O/fill_$cfg.ncu-rep $O/r2j_fill_kernel_$cfg.txt "round 2 final: fill_kernel $cfg (c2 hex27 NS 64^3 | c3 hex27 NS+T 48^3 | c5 hex8 PSPG+T+2Y 64^3 | c4 hex27 ALE 24^3), one colour launch" > /dev/null
ncu -i $O/fill_$cfg.ncu-rep --page raw --csv > $O/r2j_fill_kernel_${cfg}_raw.csv 2>/dev/null
[ $cfg != c2 ] && rm -f $O/fill_$cfg.ncu-rep
done
timeout 600 ncu --set full --clock-control none -k regex:row_sum_scale -c 1 -f -o $O/row_sum_scale python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_rss.log 2>&1; tail -1 $O/ncu_rss.log
ncu -i $O/row_sum_scale.ncu-rep --page raw --csv > $O/r2j_row_sum_scale_raw.csv 2>/dev/null; rm -f $O/row_sum_scale.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/launches_c2.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
du -sh $O
