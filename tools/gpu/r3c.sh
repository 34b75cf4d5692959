#!/bin/bash
mkdir -p gpurun_out
python tools/run_chain_once.py c1 2944 3 > gpurun_out/r3c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rr_direct_kernel -s 2 -c 1 -o gpurun_out/r3c_c1 python tools/run_chain_once.py c1 2944 3 > gpurun_out/r3c_ncu.log 2>&1
tail -n 3 gpurun_out/r3c_plain.log gpurun_out/r3c_ncu.log
