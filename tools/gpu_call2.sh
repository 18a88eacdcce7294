#!/bin/bash
# round 2, call 2: TMA-fed GEMM + segment relayout: correctness, then A/B against the kernels they replace
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest_gpu_c2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu_c2.log
tail -15 gpurun_out/r02/pytest_gpu_c2.log
for cfg in "new:" "old_gemm:TOR10_GGEMM_TMA=0" "old_rel:TOR10_RELAYOUT_MODE=index" "nopdl:TOR10_PDL=0" "fixed0:TOR10_TMA_FIXED=0" "fixed2:TOR10_TMA_FIXED=2"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02/bench_c2_$name.json 2> gpurun_out/r02/bench_c2_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02/bench_c2_$name.json"))
    print("$name", "step %.1f us"%(d["ms_per_step"]*1e3), "gemm %.1f us frac %.3f"%(d["roofline"]["ms"]*1e3, d["roofline"]["frac"]), "rel %.1f us frac %.3f"%(d["roofline_relayout"]["ms"]*1e3, d["roofline_relayout"]["frac"]), "parity", d["parity"]["max_rel_err"], d["roofline"]["kernel"][:20])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r02/bench_c2_$name.err").read()[-1500:])
PY
done
