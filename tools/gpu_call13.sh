#!/bin/bash
# single GPU: final evidence -- full gpu test suite, bench line, reference arm, ncu launch list + full capture, extras
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02/pytest_gpu_c13.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02/pytest_gpu_c13.log; tail -4 gpurun_out/r02/pytest_gpu_c13.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/r02/bench_n1_c13.json 2> gpurun_out/r02/bench_n1_c13.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02/bench_n1_c13.json"))
    print("value %.0f step %.1f us | gemm %.1f us frac %.3f peak %.1f | rel %.1f us frac %.3f | e2e %.0f serial %.0f | parity %s" % (d["value"], d["ms_per_step"]*1e3, d["roofline"]["ms"]*1e3, d["roofline"]["frac"], d["roofline"]["peak"], d["roofline_relayout"]["ms"]*1e3, d["roofline_relayout"]["frac"], d["e2e"]["value"], d["e2e"]["serial_value"], d["parity"]["max_rel_err"]))
    print("itebd", json.dumps(d.get("itebd"))[:900]); print("cpu", d.get("cpu_baseline")); print("clocks", d.get("clocks"))
except Exception as e: print("parse failed", e); print(open("gpurun_out/r02/bench_n1_c13.err").read()[-1500:])
PY
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02/bench_ref_c13.json 2> gpurun_out/r02/bench_ref_c13.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02/bench_ref_c13.json
timeout 400 python tools/bench_extra.py c2 c3 f1 > gpurun_out/r02/bench_extra_c13.jsonl 2> gpurun_out/r02/bench_extra_c13.err; echo "extra rc=$?"; cut -c1-600 gpurun_out/r02/bench_extra_c13.jsonl
CMD="python bench.py --steps 2 --warmup 3 --profile-only"
$CMD > gpurun_out/r02/plain_c13.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02/ncu_launches_c13.csv $CMD > gpurun_out/r02/ncu_l13.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r02/plain_c13b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_ggemm_f64_tma|k_relayout_segments" -s 6 -c 2 -o gpurun_out/r02/prof_c13 -f $CMD > gpurun_out/r02/ncu_f13.log 2>&1
echo "full rc=$?"
