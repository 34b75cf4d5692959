#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/configs_report.py c1 c2 > gpurun_out/r3g_c1_c2.jsonl 2> gpurun_out/r3g.err
python - <<'PY'
import json
for l in open('gpurun_out/r3g_c1_c2.jsonl'):
    d=json.loads(l); print({k:d.get(k) for k in ('config','substeps','renumber','gpu_ms','reach_substeps_per_s','parity','parity_state','host_equals_dev','clamp_pattern_equal')})
PY
timeout 600 python tools/configs_report.py c3 > gpurun_out/r3g_c3.jsonl 2>> gpurun_out/r3g.err &&
ncu --set full --clock-control none --import-source on -k regex:"uh_conv|weights_kernel|transpose|uh_state" -c 12 -o gpurun_out/r3g_c3 python tools/configs_report.py c3 > gpurun_out/r3g_ncu.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r3g_c3.jsonl'):
    d=json.loads(l); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if not isinstance(v,(dict,list))})
PY
tail -n 3 gpurun_out/r3g.err gpurun_out/r3g_ncu.log
