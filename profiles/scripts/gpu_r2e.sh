#!/bin/bash
# round 2: where does the store path go?  reductions vs first-touch stores (profiling build), scatter modes
O=gpurun_out/r2e; mkdir -p $O
B="--no-e2e --no-cpu-baseline --no-extra-configs --steps 3 --warmup 3"
PL=$PWD/goma_b200/libgoma_gpu_fill_prof.so
for cfg in c2 c5; do
for dbg in 0 4 8 12 1; do GOMA_GPU_LIB=$PL GOMA_GPU_DEBUG=$dbg GOMA_GPU_CHUNK_ELEMS=-1 timeout 600 python bench.py --config $cfg $B > $O/bench_${cfg}_dbg$dbg.json 2>/dev/null; python -c "import json;d=json.load(open('$O/bench_${cfg}_dbg$dbg.json'));print('$cfg prof-lib debug $dbg',round(d['ms_per_step'],2))"; done
for sc in 0 1 2; do GOMA_GPU_CHUNK_ELEMS=-1 timeout 600 python bench.py --config $cfg --scatter $sc $B > $O/bench_${cfg}_scatter$sc.json 2>/dev/null; python -c "import json;d=json.load(open('$O/bench_${cfg}_scatter$sc.json'));print('$cfg scatter $sc',round(d['ms_per_step'],2))"; done
done
GOMA_GPU_LIB=$PL GOMA_GPU_DEBUG=4 GOMA_GPU_CHUNK_ELEMS=20000 timeout 600 python bench.py --config c5 $B > $O/bench_c5_chunk_dbg4.json 2>/dev/null; python -c "import json;d=json.load(open('$O/bench_c5_chunk_dbg4.json'));print('c5 chunk20000 debug 4',round(d['ms_per_step'],2))"
