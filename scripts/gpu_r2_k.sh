#!/bin/bash
# ncu captures for profiles/r2_*
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/k_build.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline --no-gpu-baseline > gpurun_out/k_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flow_inverse -s 5 -c 1 -f -o gpurun_out/r2_flow_b512 python scripts/prof_flow.py 512 8 > gpurun_out/k_ncu_512.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flow_inverse -s 3 -c 1 -f -o gpurun_out/r2_flow_b8192 python scripts/prof_flow.py 8192 5 > gpurun_out/k_ncu_8192.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_refine -s 5 -c 1 -f -o gpurun_out/r2_lm_refine python bench.py --mode exact --batch 2048 --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-gpu-baseline > gpurun_out/k_ncu_lm.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:sample_kernel -c 1 -f -o gpurun_out/r2_sample python scripts/time_flow.py panda__full__lp191_5.25m 8192 > gpurun_out/k_ncu_sample.log 2>&1
echo done
