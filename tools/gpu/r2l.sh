#!/bin/bash
mkdir -p gpurun_out
for C in 8 16 32 48; do
  RR_STREAM_CHUNK_ROWS=$C timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e-variants --e2e-steps 3 --no-checks > gpurun_out/r2l_chunk_$C.json 2> gpurun_out/r2l_chunk_$C.err
  python -c "
import json; d=json.load(open('gpurun_out/r2l_chunk_$C.json')); print($C, '%.4g'%d['e2e']['value'])"
done
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 2 --no-checks > gpurun_out/r2l_variants.json 2> gpurun_out/r2l_variants.err
python -c "
import json; d=json.load(open('gpurun_out/r2l_variants.json')); print('auto %.4g'%d['e2e']['value']); print(d['e2e']['variants'].get('router_files')); print(d['e2e'].get('pcie_ceiling'))"
tail -n 3 gpurun_out/r2l_variants.err
