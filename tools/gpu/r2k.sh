#!/bin/bash
mkdir -p gpurun_out
python tools/pcie_ceiling_probe.py > gpurun_out/r2k_pcie_n1.json 2> gpurun_out/r2k_pcie_n1.err; cat gpurun_out/r2k_pcie_n1.json | cut -c1-700
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
tail -n 5 gpurun_out/r2k_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench.json'))
print('value %.4g ms %.3f'%(d['value'], d['ms_per_step']))
r=d['roofline']; print({k:r[k] for k in ('frac','frac_dram','frac_min','kernel_ms','tile_rows','units_per_launch')}, r['step'])
e=d['e2e']; print('e2e %.4g'%e['value'], e.get('pcie_ceiling'))
for k,v in e['variants'].items(): print(k, {a:(('%.4g'%b) if isinstance(b,float) else b) for a,b in v.items() if a!='api'})
print(d['checks']); print(d.get('cpu_baseline'))
PY
