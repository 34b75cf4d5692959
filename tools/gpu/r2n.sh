#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/configs_report.py c4 c5 > gpurun_out/r2n_configs.jsonl 2> gpurun_out/r2n_configs.err; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2n_configs.jsonl'):
    d=json.loads(l)
    print({k:v for k,v in d.items() if k not in ('plan',)})
PY
tail -n 6 gpurun_out/r2n_configs.err
