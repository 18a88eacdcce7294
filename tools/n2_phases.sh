#!/bin/bash
# stream-K phases on a sharded rank: full-size check with the default, then the bench for 1, 2, 3 phases
mkdir -p gpurun_out
N=${N:-2}
REPS=10 TRANSPORTS=fused timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  tools/sharded_check.py 2> gpurun_out/sharded_check.err | tee gpurun_out/sharded_check_n${N}_phases.json | tail -1
for ph in ${PHASES:-2 1 3}; do
  TOR10_SK_PHASES=$ph timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}_ph$ph.json 2> gpurun_out/bench_n${N}_ph$ph.err
  echo "phases=$ph rc=$?"; tail -1 gpurun_out/bench_n${N}_ph$ph.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'ms': d['ms_per_step'], 'value': d['value'], 'cfg': d['config'].get('tile_cfg'), 'sharded': d['config'].get('value_if_result_stays_sharded_gflops'), 'gemm_ms': d['roofline']['ms'], 'e2e': d['e2e']['value']}))"
done
