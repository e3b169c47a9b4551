#!/bin/bash
# r2y: final state of round 2 -- GPU suite, smoke, the default bench line, the reference arm
O=gpurun_out/r2y; mkdir -p $O
( timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log )
tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 200 $O/bench_default.json; tail -2 $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 200 $O/bench_reference.json
