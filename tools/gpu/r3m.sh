#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3m_smoke.log 2>&1; tail -n 2 gpurun_out/r3m_smoke.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-variants --e2e-steps 1 --e2e-rows 32 --no-checks"
$B > gpurun_out/r3m_plain.json 2> gpurun_out/r3m_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rr_|stage_|permute_|finish_|fill_' -c 60 --csv --log-file gpurun_out/r3m_launches.csv $B > gpurun_out/r3m_ncu_list.log 2>&1
$B > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'rr_direct|stage_in|stage_out' -s 9 -c 3 -o gpurun_out/r3m_c4 $B > gpurun_out/r3m_ncu_full.log 2>&1
tail -n 2 gpurun_out/r3m_ncu_full.log
