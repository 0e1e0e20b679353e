#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/am_build.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/am_launches_exact.csv python bench.py --mode exact --batch 2048 --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-gpu-baseline > gpurun_out/am_ncu_exact.log 2>&1
echo "rc $?" >> gpurun_out/am_ncu_exact.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_inverse -s 3 -c 1 -f -o gpurun_out/r2b_flow_pp_b2048 python scripts/prof_flow.py 2048 5 > gpurun_out/am_ncu_2048.log 2>&1
echo "rc $?" >> gpurun_out/am_ncu_2048.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_inverse -s 3 -c 1 -f -o gpurun_out/r2b_flow_pp_b1024 python scripts/prof_flow.py 1024 5 > gpurun_out/am_ncu_1024.log 2>&1
echo "rc $?" >> gpurun_out/am_ncu_1024.log
echo done
