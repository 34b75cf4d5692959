#!/bin/bash
mkdir -p gpurun_out
for MI in 3 4; do echo "== max_in $MI"; timeout 300 python tools/debug/confluence_debug.py $MI 2>&1 | grep -v "^  reach" | head -n 12; done
: > gpurun_out/r3a_trace.jsonl
run() { timeout 300 python tools/trace_chain.py "$@" >> gpurun_out/r3a_trace.jsonl 2>> gpurun_out/r3a.err; }
run c1 2944 1 0 0
RR_PROG_SPIN_NS=0 run c1 2944 1 0 0
RR_PROG_SPIN_NS=0 RR_GRID_CTAS=148 run c1 2944 1 0 0
python - <<'PY'
import json
for l in open('gpurun_out/r3a_trace.jsonl'):
    d=json.loads(l)
    print({k:d[k] for k in ('network','T','K','span_us','per level: done(l,g) - done(l-1,g)','per group inside a tile, chain','per group across a tile boundary, chain','per group: level-0 blocks','wait for dependencies (pre -> seen)','polls per group inside a tile, chain (median, p90)','us per poll','first row -> stores issued (15 rows + stores)')})
PY
tail -n 5 gpurun_out/r3a.err
for S in 32 0; do RR_PROG_SPIN_NS=$S timeout 300 python tools/profile_chain.py c1; done
RR_PROG_SPIN_NS=0 timeout 300 python tools/profile_chain.py c2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -x -q -m gpu > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3a_pytest.log
tail -n 6 gpurun_out/r3a_pytest.log
