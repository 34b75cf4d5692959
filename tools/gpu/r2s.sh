#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2s_chain.jsonl
for NET in c1 c2; do
  for SPIN in 32 0 200; do
    RR_PROG_SPIN_NS=$SPIN timeout 300 python tools/profile_chain.py $NET >> gpurun_out/r2s_chain.jsonl 2>> gpurun_out/r2s_chain.err
  done
  RR_NARROW_BLOCKS=0 timeout 300 python tools/profile_chain.py $NET >> gpurun_out/r2s_chain.jsonl 2>> gpurun_out/r2s_chain.err
done
cat gpurun_out/r2s_chain.jsonl; tail -n 3 gpurun_out/r2s_chain.err
