#!/bin/bash
# round 2, call 1: baseline state on a fresh box -- gpu tests, bench line with parity
mkdir -p gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r02/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest_gpu_c1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu_c1.log
tail -3 gpurun_out/r02/pytest_gpu_c1.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r02/bench_n1_c1.json 2> gpurun_out/r02/bench_n1_c1.err; echo "bench rc=$?"
cat gpurun_out/r02/bench_n1_c1.json | head -c 3000
