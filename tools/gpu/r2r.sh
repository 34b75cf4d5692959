#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_c1.py 12 > gpurun_out/r2r_c1k12.jsonl 2> gpurun_out/r2r_c1k12.err; cat gpurun_out/r2r_c1k12.jsonl; tail -n 3 gpurun_out/r2r_c1k12.err
timeout 600 python tools/profile_c1.py 1 > gpurun_out/r2r_c1k1.jsonl 2>> gpurun_out/r2r_c1k12.err; cat gpurun_out/r2r_c1k1.jsonl
