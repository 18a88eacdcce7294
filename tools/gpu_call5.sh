#!/bin/bash
# 2 GPUs: multi-GPU tests (incl. resident chaining), NVLink probe, bench N=2
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02/pytest_multi_c5.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02/pytest_multi_c5.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nvlink_probe.py > gpurun_out/r02/nvlink_peak_n2.json 2> gpurun_out/r02/nvlink_n2.err; echo "probe rc=$?"; tail -3 gpurun_out/r02/nvlink_n2.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02/nvlink_peak_n2.json"))
    for c in d["cases"]: print("%6.1f MB %-28s %7.1f us %7.1f GB/s" % (c["buffer_MB"], c["op"], c["us"], c["GBps_per_gpu_per_direction"]))
except Exception as e: print("probe parse failed", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r02/bench_n2_c5.json 2> gpurun_out/r02/bench_n2_c5.err; echo "bench rc=$?"; tail -5 gpurun_out/r02/bench_n2_c5.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02/bench_n2_c5.json").read().strip().splitlines()[-1])
    print("N=2 value %.0f GF/s step %.1f us transport" % (d["value"], d["ms_per_step"]*1e3), d["config"]["parallelism"][:60])
    print("others", json.dumps(d["config"]["other_transports"])[:900])
    print("parity", d["parity"]); print("roofline", d["roofline"]["ms"], d["roofline"]["frac"], d["roofline"]["rank_share_of_flops"])
    print("e2e", d["e2e"]["value"], d["e2e"]["serial_value"])
except Exception as e: print("bench parse failed", e)
PY
