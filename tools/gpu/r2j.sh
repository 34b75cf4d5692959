#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -8 gpurun_out/r2j_pytest.log
timeout 900 python tools/configs_report.py c2 c3 c3_pipeline > gpurun_out/r2j_configs.jsonl 2> gpurun_out/r2j_configs.err; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2j_configs.jsonl'):
    d=json.loads(l)
    print({k:v for k,v in d.items() if k in ('config','stage','renumber','gpu_ms','reach_steps_per_s','parity','wall_s','river_steps_per_s','basin_steps_per_s')})
PY
tail -n 5 gpurun_out/r2j_configs.err
