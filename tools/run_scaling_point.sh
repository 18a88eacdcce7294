#!/bin/bash
# One point of the scaling record on an N-GPU box: bench.py under torchrun (all transports, parity) + C4 / C5.
#   gpurun --gpus N -- bash tools/run_scaling_point.sh N
N=$1
mkdir -p gpurun_out/r02
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N --steps 100 --warmup 5 2> gpurun_out/r02/bench_n${N}_final.err | grep '^{' > gpurun_out/r02/bench_n${N}_final.json; echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02/bench_n${N}_final.json").read().strip().splitlines()[-1])
    print("N=$N value %.0f GF/s step %.1f us" % (d["value"], d["ms_per_step"]*1e3), "parity", d["parity"]["max_rel_err"], "| gemm share %.1f us frac %.3f | rel %.1f us | e2e %.0f serial %.0f" % (d["roofline"]["ms"]*1e3, d["roofline"]["frac"], d["roofline_relayout"]["ms"]*1e3, d["e2e"]["value"], d["e2e"]["serial_value"]))
    for k,v in d["config"]["other_transports"].items(): print("   ", k, ("%.1f us %.0f GF/s parity %s" % (v["ms_per_step"]*1e3, v["value"], v["parity"]["max_rel_err"])) if "ms_per_step" in v else v)
except Exception as e: print("parse failed", e); print(open("gpurun_out/r02/bench_n${N}_final.err").read()[-1500:])
PY
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N tools/bench_extra.py c4 c5 2> gpurun_out/r02/bench_extra_n${N}.err | grep "^{" > gpurun_out/r02/bench_extra_n${N}.jsonl; echo "extra rc=$?"; cut -c1-420 gpurun_out/r02/bench_extra_n${N}.jsonl; grep -v "^\*\|OMP" gpurun_out/r02/bench_extra_n${N}.err | grep -i "error" | tail -3 | cut -c1-300
