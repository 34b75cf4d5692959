#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; nvidia-smi topo -m 2>/dev/null | head -12
: > gpurun_out/r2m_pcie.jsonl
python tools/pcie_ceiling_probe.py --mb 1024 --reps 4 >> gpurun_out/r2m_pcie.jsonl 2> gpurun_out/r2m_pcie.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N tools/pcie_ceiling_probe.py --mb 1024 --reps 4 >> gpurun_out/r2m_pcie.jsonl 2>> gpurun_out/r2m_pcie.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 tools/pcie_ceiling_probe.py --mb 1024 --reps 4 --bind >> gpurun_out/r2m_pcie.jsonl 2>> gpurun_out/r2m_pcie.err
python - <<'PY'
import json
for l in open('gpurun_out/r2m_pcie.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['numa_bound'], {k:round(d[k]['GBps_aggregate_per_direction'],1) for k in ('h2d_only','d2h_only','both')})
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2m_bench_n8.json 2> gpurun_out/r2m_bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_bench_n8.json'))
print('%.4g'%d['value'], d['ms_per_step'], d['roofline']['step_ms_by_kernel'], 'e2e %.4g'%d['e2e']['value'])
print({k:v for k,v in d['checks'].items() if 'parity' in k or 'bitwise' in k})
PY
tail -n 4 gpurun_out/r2m_bench_n8.err
