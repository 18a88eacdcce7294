#!/bin/bash
# 8-GPU box: scaling picture (bench at N=8,4,2 with all transports + parity), resident breakdown at N=8, NVLink probe at N=8
mkdir -p gpurun_out/r02
for n in 8 4 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 50 --warmup 5 2> gpurun_out/r02/bench_n${n}_c7.err | grep '^{' > gpurun_out/r02/bench_n${n}_c7.json; echo "bench N=$n rc=$?"; tail -3 gpurun_out/r02/bench_n${n}_c7.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02/bench_n${n}_c7.json").read().strip().splitlines()[-1])
    print("N=$n value %.0f GF/s step %.1f us" % (d["value"], d["ms_per_step"]*1e3), "parity", d["parity"]["max_rel_err"], "| gemm share %.1f us frac %.3f | rel %.1f us" % (d["roofline"]["ms"]*1e3, d["roofline"]["frac"], d["roofline_relayout"]["ms"]*1e3))
    for k,v in d["config"]["other_transports"].items(): print("   ", k, ("%.1f us %.0f GF/s parity %s" % (v["ms_per_step"]*1e3, v["value"], v["parity"]["max_rel_err"])) if "ms_per_step" in v else v)
    print("    e2e", round(d["e2e"]["value"]), round(d["e2e"]["serial_value"]))
except Exception as e: print("parse failed", e)
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/resident_breakdown.py 2> gpurun_out/r02/breakdown_n8.err | grep '^{' > gpurun_out/r02/breakdown_n8_resident.json; cat gpurun_out/r02/breakdown_n8_resident.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/nvlink_probe.py 2> gpurun_out/r02/nvlink_n8.err | grep '^{' > gpurun_out/r02/nvlink_peak_n8.json
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02/nvlink_peak_n8.json").read().strip().splitlines()[-1])
    for c in d["cases"]: print("%6.1f MB %-28s %7.1f us %7.1f GB/s" % (c["buffer_MB"], c["op"], c["us"], c["GBps_per_gpu_per_direction"]))
except Exception as e: print("probe parse failed", e)
PY
