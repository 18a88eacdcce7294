#!/bin/bash
mkdir -p gpurun_out/r02
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 tools/resident_breakdown.py 2> gpurun_out/r02/breakdown_n4.err | grep '^{' > gpurun_out/r02/breakdown_n4_resident.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/breakdown_n4_resident.json").read().strip().splitlines()[-1])
tl=d.pop("timeline",None); print(d)
for e in tl or []: print("  %-62s %8.1f us  gap %s" % (e["name"], e["us"], e["gap_before_us"]))
PY
tail -3 gpurun_out/r02/breakdown_n4.err | cut -c1-300
