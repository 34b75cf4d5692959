#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python tools/sharded_router_check.py --gpus 2 > gpurun_out/r2h_sharded_router.json 2> gpurun_out/r2h_sharded_router.err; echo "sharded rc=$?"
cat gpurun_out/r2h_sharded_router.json; tail -n 5 gpurun_out/r2h_sharded_router.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench_n2.json'))
print('%.4g'%d['value'], d['ms_per_step'], d['roofline']['step_ms_by_kernel'], 'e2e %.4g'%d['e2e']['value'], d['checks'])
PY
tail -n 8 gpurun_out/r2h_bench_n2.err
