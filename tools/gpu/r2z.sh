#!/bin/bash
mkdir -p gpurun_out
for MI in 2 3 4; do
  echo "== max_in $MI"; timeout 300 python tools/debug/confluence_debug.py $MI 2>&1 | grep -v "^  reach" | head -n 12
  echo "== max_in $MI, no narrow blocks"; RR_NARROW_BLOCKS=0 timeout 300 python tools/debug/confluence_debug.py $MI 2>&1 | grep -v "^  reach" | head -n 12
done
