#!/bin/bash
# N-GPU check + bench of the sharded stream-K path (N from the environment, default 4)
mkdir -p gpurun_out
N=${N:-4}
TOR10_SHARDED_SK=1 REPS=10 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  tools/sharded_check.py 2> gpurun_out/sharded_check.err | tee gpurun_out/sharded_check_n$N.json | tail -1
for sk in 1 0; do
  TOR10_SHARDED_SK=$sk timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}_sk$sk.json 2> gpurun_out/bench_n${N}_sk$sk.err
  echo "sk=$sk rc=$?"; tail -1 gpurun_out/bench_n${N}_sk$sk.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'ms': d['ms_per_step'], 'value': d['value'], 'cfg': d['config'].get('tile_cfg'), 'sharded': d['config'].get('value_if_result_stays_sharded_gflops'), 'gemm_ms': d['roofline']['ms'], 'e2e': d['e2e']['value']}))"
done
