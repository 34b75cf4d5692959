#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2c_ref.json 2> gpurun_out/r2c_ref.err; echo "ref rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2c_ref.err gpurun_out/r2c_bench.err
python - <<'PY'
import json
for f in ('r2c_ref','r2c_bench'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, d['value'], d.get('ms_per_step'), d.get('e2e',{}).get('value'), d.get('checks'), d.get('cpu_baseline'))
        if 'roofline' in d: print({k:v for k,v in d['roofline'].items() if k.startswith('frac') or k=='step'})
    except Exception as e: print(f,'failed',e)
PY
nproc; free -g | head -2
