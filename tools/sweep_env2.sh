#!/bin/bash
# usage: tools/sweep_env2.sh "VAR1=a VAR2=b" "VAR1=c VAR2=d" ...  -> one bench summary line per setting
for v in "$@"; do
  env $v timeout 300 python bench.py --no-cpu-baseline --steps 100 2>/dev/null | VAL="$v" python -c "import os,sys,json; d=json.loads(sys.stdin.read()); print(os.environ['VAL'], round(d['value'],1), round(d['ms_per_step']*1e3,2), round(d['roofline']['ms']*1e3,2), d['config']['tile_cfg'])"
done
