#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e-variants --e2e-steps 1 --e2e-rows 32 --no-checks"
run() { # name, env, extra args
  env $2 timeout 600 $B $3 > gpurun_out/r2p_$1.json 2> gpurun_out/r2p_$1.err
  python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(f'gpurun_out/r2p_{sys.argv[1]}.json')); print(sys.argv[1], '%.4g'%d['value'], {k:round(v,3) for k,v in d['roofline']['step_ms_by_kernel'].items()})
except Exception as e: print(sys.argv[1], 'failed', e)
PY
}
run c4_base "X=1" ""
run c4_tile128 "X=1" "--time-tile 128"
run c4_narrow1k "RR_NARROW_BLOCKS=1024" ""
run c4_narrow0 "RR_NARROW_BLOCKS=0" ""
run c4_narrow32k "RR_NARROW_BLOCKS=32768" ""
N8="--reaches 875000 --basins 625 --rows 1920"
run n8_base "X=1" "$N8"
run n8_narrow1k "RR_NARROW_BLOCKS=1024" "$N8"
run n8_narrow0 "RR_NARROW_BLOCKS=0" "$N8"
run n8_tile128 "X=1" "$N8 --time-tile 128"
for NB in 4096 0 100000; do
RR_NARROW_BLOCKS=$NB timeout 300 python tools/configs_report.py c2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    if d.get('renumber')=='auto': print('c2 narrow $NB', d['gpu_ms'])"
done
