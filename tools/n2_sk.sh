#!/bin/bash
# 2-GPU check of the stream-K + fused all-gather kernel: parity test, then the bench with and without it
mkdir -p gpurun_out
N=${N:-2}
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/multi_sk_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/multi_sk_pytest.log
tail -5 gpurun_out/multi_sk_pytest.log
TOR10_SHARDED_SK=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  tools/sharded_check.py 2> gpurun_out/sharded_check.err | tee gpurun_out/sharded_check_n$N.json | cut -c1-700
for sk in 1 0; do
  TOR10_SHARDED_SK=$sk timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}_sk$sk.json 2> gpurun_out/bench_n${N}_sk$sk.err
  echo "sk=$sk rc=$?"; tail -1 gpurun_out/bench_n${N}_sk$sk.json | cut -c1-600
done
