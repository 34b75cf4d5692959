#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_chain.py c1; timeout 300 python tools/profile_chain.py c2
timeout 600 python tools/profile_c1.py 1 2> gpurun_out/r3f.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['substeps'], d['staging'], d['time_tile'], d['tile_stride'], d['tile_rows'], d['ms_by_class_last_rep'], d['checksum'])"
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r3f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3f_pytest.log
tail -n 6 gpurun_out/r3f_pytest.log
