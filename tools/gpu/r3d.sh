#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r3d_trace.jsonl
run() { timeout 300 python tools/trace_chain.py "$@" >> gpurun_out/r3d_trace.jsonl 2>> gpurun_out/r3d.err; }
run c1 2944 1 0 0 save
RR_PROG_SPIN_NS=0 run c1 2944 1 0 0
python - <<'PY'
import json
for l in open('gpurun_out/r3d_trace.jsonl'):
    d=json.loads(l)
    print({k:d[k] for k in ('network','T','K','span_us','per level: done(l,g) - done(l-1,g)','per group inside a tile, chain','per group across a tile boundary, chain','per group: level-0 blocks','wait for dependencies (pre -> seen)','polls per group inside a tile, chain (median, p90)','us per poll','first row -> stores issued (15 rows + stores)')})
PY
tail -n 5 gpurun_out/r3d.err
for S in 32 0; do RR_PROG_SPIN_NS=$S timeout 300 python tools/profile_chain.py c1; done
for S in 32 0; do RR_PROG_SPIN_NS=$S timeout 300 python tools/profile_chain.py c2; done
timeout 300 python tools/debug/confluence_debug.py 3 2>&1 | grep -v "^  reach" | head -n 6
