#!/bin/bash
# 2 GPUs: multi tests (incl. sharded SVD chain), C4/C5 at N=1 and N=2
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02/pytest_multi_c12.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|rank [0-9]" gpurun_out/r02/pytest_multi_c12.log | head -10 | cut -c1-300
timeout 400 python tools/bench_extra.py c4 c5 > gpurun_out/r02/bench_extra_n1.jsonl 2> gpurun_out/r02/bench_extra_n1.err; echo "extra n1 rc=$?"; cut -c1-700 gpurun_out/r02/bench_extra_n1.jsonl; tail -3 gpurun_out/r02/bench_extra_n1.err | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 tools/bench_extra.py c4 c5 2> gpurun_out/r02/bench_extra_n2.err | grep '^{' > gpurun_out/r02/bench_extra_n2.jsonl; echo "extra n2 rc=$?"; cut -c1-700 gpurun_out/r02/bench_extra_n2.jsonl; grep -v "^\*\|OMP" gpurun_out/r02/bench_extra_n2.err | tail -5 | cut -c1-300
