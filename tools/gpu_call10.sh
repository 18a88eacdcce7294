#!/bin/bash
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_svd.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02/pytest_c10.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02/pytest_c10.log | cut -c1-250
timeout 300 python tools/itebd_bench.py 32 1000 > gpurun_out/r02/itebd_bench.json 2> gpurun_out/r02/itebd_bench.err; echo "itebd rc=$?"; cat gpurun_out/r02/itebd_bench.json; tail -5 gpurun_out/r02/itebd_bench.err | cut -c1-250
python tools/profile_itebd.py > gpurun_out/r02/itebd_host_profile2.txt 2>&1; head -40 gpurun_out/r02/itebd_host_profile2.txt | cut -c1-160
