#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -12 gpurun_out/r2q_pytest.log
timeout 600 python tools/configs_report.py c1 > gpurun_out/r2q_c1.jsonl 2> gpurun_out/r2q_c1.err
python - <<'PY'
import json
for l in open('gpurun_out/r2q_c1.jsonl'):
    d=json.loads(l); print({k:d[k] for k in ('config','substeps','gpu_ms','reach_substeps_per_s','parity','parity_state','host_equals_dev','clamp_pattern_equal')})
PY
