#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_chain.py c1 512 1 > gpurun_out/r2u_trace_c1.json 2> gpurun_out/r2u.err
timeout 300 python tools/trace_chain.py c1 40 12 > gpurun_out/r2u_trace_c1_k12.json 2>> gpurun_out/r2u.err
cat gpurun_out/r2u_trace_c1.json; tail -n 5 gpurun_out/r2u.err
