#!/bin/bash
# r2r: final state of round 2 -- GPU suite, smoke, the default bench line, the reference arm, one ncu capture of the
# 4-rows-per-CTA row-sum scaling kernel (DRAM bytes against 16 * nnz)
O=gpurun_out/r2r; mkdir -p $O
( timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log )
tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json; tail -3 $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 300 $O/bench_reference.json
timeout 600 ncu --set full --clock-control none -k regex:row_sum_scale -c 1 -f -o $O/row_sum_scale python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_rss.log 2>&1; tail -1 $O/ncu_rss.log
ncu -i $O/row_sum_scale.ncu-rep --page raw --csv > $O/r2r_row_sum_scale_raw.csv 2>/dev/null; rm -f $O/row_sum_scale.ncu-rep
du -sh $O
