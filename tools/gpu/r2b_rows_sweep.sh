#!/bin/bash
# route-kernel time vs rows per launch: the fixed part is the wavefront's fill + drain (critical path of the last tile)
mkdir -p gpurun_out
for R in 64 128 240 480; do
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e-variants --e2e-steps 1 --e2e-rows 16 --rows $R > gpurun_out/r2b_rows_$R.json 2> gpurun_out/r2b_rows_$R.err
done
python - <<'PY'
import json
for R in (64,128,240,480):
    try:
        d=json.load(open(f'gpurun_out/r2b_rows_{R}.json'))
        print(R, d['value'], d['roofline']['step_ms_by_kernel'])
    except Exception as e: print(R, 'failed', e)
PY
