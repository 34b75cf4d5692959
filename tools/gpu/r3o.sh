#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r3o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3o_pytest.log
tail -n 3 gpurun_out/r3o_pytest.log
timeout 600 python bench.py > gpurun_out/r3o_bench_n1.json 2> gpurun_out/r3o_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3o_bench_n1.json')); r=d['roofline']
print('value %.4g ms/step %.3f e2e %.4g launches %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches']))
print({k:round(v,3) for k,v in r['step_ms_by_kernel'].items()}, 'frac',round(r['frac'],3),'frac_dram',round(r['frac_dram'],3),'frac_min',round(r['frac_min'],3), 'step', {k:(round(v,3) if isinstance(v,float) else v) for k,v in r['step'].items()})
print('parity_ok', d['checks']['parity_ok'], 'clocks', d.get('clocks'))
PY
