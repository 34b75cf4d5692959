#!/bin/bash
# Round 2, call A: ncu evidence for the kernel that ships (direct-exchange rr_wavefront_kernel<1>) and both permutes.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-variants --e2e-steps 1"
N8="$B --reaches 875000 --basins 625 --rows 1920"
$B > gpurun_out/r2a_plain.json 2> gpurun_out/r2a_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2a_launches.csv $B > gpurun_out/r2a_ncu_list.log 2>&1
$B > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'rr_wavefront|permute_to' -s 15 -c 5 -o gpurun_out/r2a_c4 $B > gpurun_out/r2a_ncu_full.log 2>&1
$N8 > gpurun_out/r2a_plain_n8shape.json 2> gpurun_out/r2a_plain_n8shape.err &&
ncu --set full --clock-control none --import-source on -k regex:'rr_wavefront' -s 3 -c 1 -o gpurun_out/r2a_n8shape $N8 > gpurun_out/r2a_ncu_n8shape.log 2>&1
ls -la gpurun_out | tail -20
