#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2w_trace.jsonl
run() { timeout 300 python tools/trace_chain.py "$@" >> gpurun_out/r2w_trace.jsonl 2>> gpurun_out/r2w.err; }
RR_UNSAFE_NOFENCE=1 run c1 2944 1 0 0
RR_UNSAFE_NOFENCE=1 run c1 2944 1 2944 0
RR_UNSAFE_NOFENCE=1 run c1 2944 1 256 0
RR_GRID_CTAS=210 run c1 2944 1 2944 0
RR_UNSAFE_NOFENCE=1 RR_GRID_CTAS=210 run c1 2944 1 2944 0
python - <<'PY'
import json
for l in open('gpurun_out/r2w_trace.jsonl'):
    d=json.loads(l)
    print({k:d[k] for k in ('T','K','time_tile','tile_stride','tile_rows','gpt','span_us','per level: done(l,g) - done(l-1,g)','per group inside a tile, chain','per group across a tile boundary, chain','per group: level-0 blocks','stores issued -> release returned','seen -> first row (data load + 1 row)','first row -> stores issued (15 rows + stores)','hop_level_us: upstream release returned -> dependencies seen')})
PY
tail -n 5 gpurun_out/r2w.err
