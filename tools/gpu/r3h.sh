#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug/confluence_debug.py 3 2>&1 | grep -v "^  reach" | head -n 6
timeout 300 python tools/profile_chain.py c1; timeout 300 python tools/profile_chain.py c2
timeout 900 python tools/configs_report.py c1 c2 2> gpurun_out/r3h.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d.get(k) for k in ('config','substeps','renumber','gpu_ms','parity','parity_state','host_equals_dev','clamp_pattern_equal')})"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -x -q -m gpu > gpurun_out/r3h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3h_pytest.log
tail -n 4 gpurun_out/r3h_pytest.log
ncu --set full --clock-control none --import-source on -k regex:"uh_conv" -c 2 -o gpurun_out/r3h_uh python tools/configs_report.py c3 > gpurun_out/r3h_ncu.log 2>&1; tail -n 2 gpurun_out/r3h_ncu.log
timeout 600 python -m pytest tests/test_gpu_transforms.py tests/test_gpu_stream.py -x -q -m gpu 2>&1 | tail -n 3
timeout 600 python tools/configs_report.py c3 2>> gpurun_out/r3h.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if not isinstance(v,(dict,list))})"
