#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e-variants --e2e-steps 1 --e2e-rows 32 --no-checks"
run() { # name, env, extra
  env $2 timeout 600 $B $3 > gpurun_out/r3n_$1.json 2> gpurun_out/r3n_$1.err
  python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(f'gpurun_out/r3n_{sys.argv[1]}.json')); print(sys.argv[1], '%.4g'%d['value'], {k:round(v,3) for k,v in d['roofline']['step_ms_by_kernel'].items()}, 'renumbered', d['config'].get('plan',{}).get('renumbered'))
except Exception as e: print(sys.argv[1], 'failed', e)
PY
}
run order_level "X=1" "--order level"
run half_occupancy "RR_GRID_CTAS=148" ""
