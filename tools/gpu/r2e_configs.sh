#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/configs_report.py c1 c2 c3 c3_pipeline > gpurun_out/r2e_configs.jsonl 2> gpurun_out/r2e_configs.err; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2e_configs.jsonl'):
    d=json.loads(l)
    keep={k:v for k,v in d.items() if k not in ('plan',) and not isinstance(v,(list,dict))}
    print(keep)
PY
tail -n 5 gpurun_out/r2e_configs.err
