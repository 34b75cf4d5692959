#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2v_trace.jsonl
run() { timeout 300 python tools/trace_chain.py "$@" >> gpurun_out/r2v_trace.jsonl 2>> gpurun_out/r2v.err; }
run c1 2944 1 0 0
run c1 2944 1 256 0 save
run c1 2944 1 2944 0 save
run c1 2944 1 0 6
run c1 2944 1 0 8
run c1 2944 1 256 20
run c1 240 12 0 0
run c1 240 12 2880 0
python - <<'PY'
import json
for l in open('gpurun_out/r2v_trace.jsonl'):
    d=json.loads(l)
    print({k:d[k] for k in ('T','K','time_tile','tile_stride','tile_rows','gpt','span_us','per level: done(l,g) - done(l-1,g)','per group inside a tile, chain','per group across a tile boundary, chain','per group: level-0 blocks')})
PY
tail -n 5 gpurun_out/r2v.err
