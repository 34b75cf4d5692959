#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e-variants --e2e-steps 1 --e2e-rows 32 --no-checks"
run() { # name, env
  env $2 timeout 600 $B > gpurun_out/r3j_$1.json 2> gpurun_out/r3j_$1.err
  python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(f'gpurun_out/r3j_{sys.argv[1]}.json')); print(sys.argv[1], '%.4g'%d['value'], {k:round(v,3) for k,v in d['roofline']['step_ms_by_kernel'].items()})
except Exception as e: print(sys.argv[1], 'failed', e)
PY
}
run spin0 "X=1"
run spin64 "RR_PROG_SPIN_NS=64"


run narrow512 "RR_NARROW_BLOCKS=512"
run narrow1024 "RR_NARROW_BLOCKS=1024"
run narrow2048 "RR_NARROW_BLOCKS=2048"
for NB in 1024 4096; do RR_NARROW_BLOCKS=$NB timeout 300 python tools/profile_chain.py c2 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c2 narrow $NB', {k:d[k] for k in ('T64_route_ms','T1024_route_ms','T2944_route_ms')})"; done
