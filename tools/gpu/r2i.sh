#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -6 gpurun_out/r2i_pytest.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-variants --e2e-steps 1 --e2e-rows 32 --no-checks"
$B > gpurun_out/r2i_plain.json 2> gpurun_out/r2i_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rr_|stage_|permute_|finish_' -c 60 --csv --log-file gpurun_out/r2i_launches.csv $B > gpurun_out/r2i_ncu_list.log 2>&1
$B > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'rr_direct|stage_in|stage_out' -s 9 -c 3 -o gpurun_out/r2i_c4 $B > gpurun_out/r2i_ncu_full.log 2>&1
N8="$B --reaches 875000 --basins 625 --rows 1920"
$N8 > gpurun_out/r2i_plain_n8.json 2> gpurun_out/r2i_plain_n8.err &&
ncu --set full --clock-control none --import-source on -k regex:'rr_direct' -s 3 -c 1 -o gpurun_out/r2i_n8shape $N8 > gpurun_out/r2i_ncu_n8.log 2>&1
ls -la gpurun_out | grep r2i
