#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3l_pytest.log
tail -n 4 gpurun_out/r3l_pytest.log
timeout 900 python bench.py > gpurun_out/r3l_bench_n1.json 2> gpurun_out/r3l_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r3l_bench_n1.json'))
    r=d['roofline']
    print('value %.4g ms/step %.3f e2e %.4g launches %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches']))
    print({k:round(v,3) for k,v in r['step_ms_by_kernel'].items()}, 'frac',round(r['frac'],3),'frac_dram',r.get('frac_dram'),'frac_min',round(r['frac_min'],3))
    print('parity_ok', d.get('checks',{}).get('parity_ok'))
except Exception as e: print('bench parse failed', e)
PY
timeout 600 python tools/configs_report.py c1 c2 > gpurun_out/r3l_c1_c2.jsonl 2> gpurun_out/r3l.err
python - <<'PY'
import json
for l in open('gpurun_out/r3l_c1_c2.jsonl'):
    d=json.loads(l); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in ('config','substeps','renumber','gpu_ms','reach_substeps_per_s','reach_steps_per_s','parity')})
PY
