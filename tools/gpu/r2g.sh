#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/configs_report.py c1 c2 > gpurun_out/r2g_configs.jsonl 2> gpurun_out/r2g_configs.err; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2g_configs.jsonl'):
    d=json.loads(l)
    print({k:v for k,v in d.items() if k in ('config','substeps','renumber','gpu_ms','reach_substeps_per_s','reach_steps_per_s','parity')})
PY
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e-variants --e2e-steps 1 --e2e-rows 32 --no-checks"
timeout 600 $B > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
timeout 600 $B --reaches 875000 --basins 625 --rows 1920 > gpurun_out/r2g_n8shape.json 2> gpurun_out/r2g_n8shape.err; echo "n8 rc=$?"
timeout 600 $B --tile-stride 1 > gpurun_out/r2g_bench_d1.json 2> gpurun_out/r2g_bench_d1.err; echo "bench d1 rc=$?"
python - <<'PY'
import json
for f in ('r2g_bench','r2g_n8shape','r2g_bench_d1'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f,'%.4g'%d['value'], d['ms_per_step'], d['roofline']['step_ms_by_kernel'])
PY
timeout 900 python -m pytest tests/test_gpu_stress.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; tail -3 gpurun_out/r2g_pytest.log
