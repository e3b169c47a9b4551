#!/bin/bash
# r2u: compute-sanitizer over the small cases of every kernel family: memcheck (out-of-bounds / misaligned accesses),
# racecheck (shared-memory hazards), initcheck (reads of uninitialised device memory)
O=gpurun_out/r2u; mkdir -p $O
KM='fixture or tiny or irregular or ghost or q2p1 or zero_row or async or row_sum or csr or host_stream or detJ or global_h'
KR='fixture and first_touch or row_sum_scaling_matches_oracle or host_stream_chunks and c3_hex27 or csr_layout and first_touch and (c3_hex27 or c5_hex8)'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$KM" > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/memcheck.log | tail -2
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$KR" > $O/racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/racecheck.log | tail -2
timeout 1200 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$KR" > $O/initcheck.log 2>&1; echo "initcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/initcheck.log | tail -2
