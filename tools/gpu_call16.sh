#!/bin/bash
mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -2
for cfg in "rs3:" "rs2:TOR10_RELAYOUT_SEG_CTAS=2" "rs4:TOR10_RELAYOUT_SEG_CTAS=4"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-itebd > gpurun_out/r02/bench_c16_$name.json 2> gpurun_out/r02/bench_c16_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02/bench_c16_$name.json"))
    print("$name", "step %.1f us"%(d["ms_per_step"]*1e3), "gemm %.1f us frac %.3f"%(d["roofline"]["ms"]*1e3, d["roofline"]["frac"]), "rel %.1f us frac %.3f"%(d["roofline_relayout"]["ms"]*1e3, d["roofline_relayout"]["frac"]), "parity", d["parity"]["max_rel_err"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r02/bench_c16_$name.err").read()[-800:])
PY
done
CMD="python bench.py --steps 4 --warmup 3 --profile-only"
$CMD > gpurun_out/r02/plain_c16.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_ggemm_f64_tma|k_relayout_segments|FillFunctor" -c 40 --csv --log-file gpurun_out/r02/ncu_launches_c16.csv $CMD > gpurun_out/r02/ncu_l16.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r02/plain_c16b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_ggemm_f64_tma|k_relayout_segments" -s 6 -c 2 -o gpurun_out/r02/prof_c16 -f $CMD > gpurun_out/r02/ncu_f16.log 2>&1
echo "full rc=$?"
