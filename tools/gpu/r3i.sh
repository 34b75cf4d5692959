#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3i_pytest.log
tail -n 4 gpurun_out/r3i_pytest.log
timeout 900 python bench.py > gpurun_out/r3i_bench_n1.json 2> gpurun_out/r3i_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r3i_bench_n1.json'))
    r=d['roofline']
    print('value %.4g ms/step %.3f e2e %.4g launches %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches']))
    print({k:round(v,3) for k,v in r['step_ms_by_kernel'].items()}, 'frac',round(r['frac'],3),'frac_dram',r.get('frac_dram'),'frac_min',round(r['frac_min'],3))
    print('checks', d.get('checks'))
except Exception as e: print('bench parse failed', e)
PY
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-variants --e2e-steps 1 --e2e-rows 32 --no-checks"
$B > gpurun_out/r3i_plain.json 2> gpurun_out/r3i_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rr_|stage_|permute_|finish_|fill_' -c 60 --csv --log-file gpurun_out/r3i_launches.csv $B > gpurun_out/r3i_ncu_list.log 2>&1
$B > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'rr_direct|stage_in|stage_out|fill_sentinel' -s 12 -c 4 -o gpurun_out/r3i_c4 $B > gpurun_out/r3i_ncu_full.log 2>&1
tail -n 2 gpurun_out/r3i_ncu_full.log
timeout 900 python tools/configs_report.py c1 c2 > gpurun_out/r3i_c1_c2.jsonl 2> gpurun_out/r3i.err
timeout 600 python tools/configs_report.py c3 > gpurun_out/r3i_c3.jsonl 2>> gpurun_out/r3i.err
python - <<'PY'
import json
for f in ('gpurun_out/r3i_c1_c2.jsonl','gpurun_out/r3i_c3.jsonl'):
    for l in open(f):
        d=json.loads(l); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if not isinstance(v,(dict,list)) and k in ('config','stage','substeps','renumber','gpu_ms','reach_substeps_per_s','parity','frac_of_hbm','fp64_tflops','river_steps_per_s','reach_steps_per_s')})
PY
ls -la gpurun_out | grep r3i
