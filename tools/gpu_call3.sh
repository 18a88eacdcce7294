#!/bin/bash
mkdir -p gpurun_out/r02
python tools/trace_ggemm.py > gpurun_out/r02/trace_tma.txt 2>&1; echo "trace rc=$?"; tail -30 gpurun_out/r02/trace_tma.txt
CMD="python bench.py --steps 2 --warmup 3 --profile-only"
$CMD > gpurun_out/r02/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02/ncu_launches_c3.csv $CMD > gpurun_out/r02/ncu_l.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r02/plain_c3b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_ggemm_f64_tma|k_relayout_segments" -s 6 -c 4 -o gpurun_out/r02/prof_c3 -f $CMD > gpurun_out/r02/ncu_f.log 2>&1
echo "full rc=$?"; tail -5 gpurun_out/r02/ncu_f.log; ls -la gpurun_out/r02/
