#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug/confluence_debug.py 3 > gpurun_out/r2y_debug.log 2>&1; cat gpurun_out/r2y_debug.log | tail -n 40
: > gpurun_out/r2y_trace.jsonl
run() { timeout 300 python tools/trace_chain.py "$@" >> gpurun_out/r2y_trace.jsonl 2>> gpurun_out/r2y.err; }
run c1 2944 1 0 0 save
run c1 240 12 0 0
run c2 1440 1 0 0
python - <<'PY'
import json
for l in open('gpurun_out/r2y_trace.jsonl'):
    d=json.loads(l)
    print({k:d[k] for k in ('network','T','K','time_tile','tile_rows','gpt','span_us','per level: done(l,g) - done(l-1,g)','per group inside a tile, chain','per group across a tile boundary, chain','per group: level-0 blocks','wait for dependencies (pre -> seen)','seen -> first row (data load + 1 row)','first row -> stores issued (15 rows + stores)')})
PY
tail -n 5 gpurun_out/r2y.err
timeout 300 python tools/profile_chain.py c1; timeout 300 python tools/profile_chain.py c2
