#!/bin/bash
# r2v2: 2 GPUs, final library (hybrid element hand-out): the weak-scaling line as the driver launches it, device-timed part only
O=gpurun_out/r2v2; mkdir -p $O
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/weak_c2_n2.json 2> $O/weak_c2_n2.err
python - <<'PY'
import json
d=None
for l in open("gpurun_out/r2v2/weak_c2_n2.json"):
    if l.startswith("{"): d=json.loads(l)
print(d["n_gpus"], d["scaling"], round(d["value"]/1e6,3), "M el/s", round(d["ms_per_step"],2), "ms", d["config"]["halo"][:60] if d["config"].get("halo") else None)
PY
tail -2 $O/weak_c2_n2.err
