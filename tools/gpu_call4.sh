#!/bin/bash
mkdir -p gpurun_out/r02
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02/bench_c4_$name.json 2> gpurun_out/r02/bench_c4_$name.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02/bench_c4_$name.json"))
    print("$name", "step %.1f us"%(d["ms_per_step"]*1e3), "gemm %.1f us frac %.3f"%(d["roofline"]["ms"]*1e3, d["roofline"]["frac"]), "rel %.1f us frac %.3f"%(d["roofline_relayout"]["ms"]*1e3, d["roofline_relayout"]["frac"]), "parity", d["parity"]["max_rel_err"])
except Exception as e:
    print("$name failed rc=$rc", e); print(open("gpurun_out/r02/bench_c4_$name.err").read()[-1500:])
PY
}
run base A=1
run ctas3 TOR10_TMA_CTAS=3
run ctas3_f2 TOR10_TMA_CTAS=3 TOR10_TMA_FIXED=2
run ctas3_f05 TOR10_TMA_CTAS=3 TOR10_TMA_FIXED=0.5
run f15 TOR10_TMA_FIXED=1.5
run mp1 TOR10_SK_MINPIECE=1
run mp4 TOR10_SK_MINPIECE=4
run rs2 TOR10_RELAYOUT_SEG_CTAS=2
run rs4 TOR10_RELAYOUT_SEG_CTAS=4
run rs8 TOR10_RELAYOUT_SEG_CTAS=8
