#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -15 gpurun_out/r2d_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e-variants --e2e-steps 1 --e2e-rows 32"
for S in auto direct-nohw; do
  timeout 600 $B --staging $S > gpurun_out/r2d_bench_$S.json 2> gpurun_out/r2d_bench_$S.err; echo "bench $S rc=$?"
done
timeout 600 $B --reaches 875000 --basins 625 --rows 1920 > gpurun_out/r2d_n8shape.json 2> gpurun_out/r2d_n8shape.err; echo "n8shape rc=$?"
python - <<'PY'
import json
for f in ('r2d_bench_auto','r2d_bench_direct-nohw','r2d_n8shape'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, '%.4g'%d['value'], d.get('ms_per_step'), d['roofline']['step_ms_by_kernel'], d.get('checks',{}).get('parity_ok'), d.get('checks',{}).get('oracle_parity_max_rel'))
    except Exception as e: print(f,'failed',e)
PY
tail -5 gpurun_out/r2d_*.err
