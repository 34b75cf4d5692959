#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_transforms.py tests/test_gpu_stream.py tests/test_routers_gpu.py -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; tail -3 gpurun_out/r2o_pytest.log
for V in smem reg; do
  if [ $V = reg ]; then export RR_UH_REGISTERS=1; else unset RR_UH_REGISTERS; fi
  timeout 600 python tools/configs_report.py c3 > gpurun_out/r2o_c3_$V.jsonl 2> gpurun_out/r2o_c3_$V.err
  python - "$V" <<'PY'
import json,sys
for l in open(f'gpurun_out/r2o_c3_{sys.argv[1]}.jsonl'):
    d=json.loads(l)
    if d.get('stage')=='uh_convolve': print(sys.argv[1], {k:d[k] for k in ('gpu_ms','fp64_tflops','frac_of_hbm','parity_subset','parity_state_subset')})
PY
done
