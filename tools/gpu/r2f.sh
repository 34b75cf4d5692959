#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -22 gpurun_out/r2f_pytest.log
timeout 900 python tools/configs_report.py c1 c2 > gpurun_out/r2f_configs.jsonl 2> gpurun_out/r2f_configs.err; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2f_configs.jsonl'):
    d=json.loads(l)
    print({k:v for k,v in d.items() if k in ('config','substeps','renumber','gpu_ms','reach_substeps_per_s','reach_steps_per_s','parity')})
PY
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e-variants --e2e-steps 2 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print('%.4g'%d['value'], d['ms_per_step'], d['roofline']['step_ms_by_kernel'], 'e2e %.4g'%d['e2e']['value'], d['checks'].get('parity_ok'))
PY
tail -n 5 gpurun_out/r2f_bench.err gpurun_out/r2f_configs.err
