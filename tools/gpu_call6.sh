#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02/pytest_multi_c6.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|error:" gpurun_out/r02/pytest_multi_c6.log | head -20
for tr in resident fused; do
TOR10_GATHER=$tr timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/resident_breakdown.py 2> gpurun_out/r02/breakdown_n2_$tr.err | grep '^{' > gpurun_out/r02/breakdown_n2_$tr.json; cat gpurun_out/r02/breakdown_n2_$tr.json; tail -3 gpurun_out/r02/breakdown_n2_$tr.err | cut -c1-300
done
