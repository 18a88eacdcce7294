#!/bin/bash
# 8-GPU box: smoke (2-rank chain), sharded_check N=8 (all transports, full size), bench N=8 and N=4
mkdir -p gpurun_out/r02
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/smoke_c11.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02/smoke_c11.log | cut -c1-200
REPS=8 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/sharded_check.py 2> gpurun_out/r02/sharded_check_n8.err | grep '^{' > gpurun_out/r02/sharded_check_n8.json; echo "sharded_check rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02/sharded_check_n8.json')); print(d['world'], d['ranks'][0]); print('all ok', all(v['bit_reproducible'] and v['max_rel_err_vs_single']<=1e-13 for r in d['ranks'] for v in r.values()))"
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n bench.py --gpus $n --steps 50 --warmup 5 2> gpurun_out/r02/bench_n${n}_c11.err | grep '^{' > gpurun_out/r02/bench_n${n}_c11.json; echo "bench N=$n rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02/bench_n${n}_c11.json").read().strip().splitlines()[-1])
    print("N=$n value %.0f GF/s step %.1f us" % (d["value"], d["ms_per_step"]*1e3), "parity", d["parity"]["max_rel_err"], "| gemm share %.1f us frac %.3f | rel %.1f us" % (d["roofline"]["ms"]*1e3, d["roofline"]["frac"], d["roofline_relayout"]["ms"]*1e3))
    for k,v in d["config"]["other_transports"].items(): print("   ", k, ("%.1f us %.0f GF/s" % (v["ms_per_step"]*1e3, v["value"])) if "ms_per_step" in v else v)
except Exception as e: print("parse failed", e)
PY
done
