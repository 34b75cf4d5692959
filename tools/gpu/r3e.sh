#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_c1.py 1 > gpurun_out/r3e_c1k1.jsonl 2> gpurun_out/r3e.err
timeout 600 python tools/profile_c1.py 12 > gpurun_out/r3e_c1k12.jsonl 2>> gpurun_out/r3e.err
python - <<'PY'
import json
for f in ('gpurun_out/r3e_c1k1.jsonl','gpurun_out/r3e_c1k12.jsonl'):
    for l in open(f):
        d=json.loads(l); print(d['substeps'], d['staging'], d['time_tile'], d['tile_stride'], d['tile_rows'], d['ms_by_class_last_rep'], d['checksum'])
PY
tail -n 3 gpurun_out/r3e.err
