#!/bin/bash
# r2w: several element blocks / materials -- the full GPU suite (new mm_* fixtures from the reference included)
O=gpurun_out/r2w; mkdir -p $O
( timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log )
tail -5 $O/pytest.log
grep -c "mm_" $O/pytest.log
python -m pytest tests -m gpu -q -k "mm_ or materials" 2>&1 | tail -3
