#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2t_chain.jsonl
for G in 296 148 74 37 18; do
  for SPIN in 32 1000; do
    RR_GRID_CTAS=$G RR_PROG_SPIN_NS=$SPIN timeout 300 python tools/profile_chain.py c1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); d['grid']=$G; print(json.dumps({k:d[k] for k in ('grid','spin_ns','T64_route_ms','T2944_route_ms','us_per_group','us_per_level_at_T64')}))" >> gpurun_out/r2t_chain.jsonl 2>> gpurun_out/r2t_chain.err
  done
done
cat gpurun_out/r2t_chain.jsonl; tail -n 3 gpurun_out/r2t_chain.err
