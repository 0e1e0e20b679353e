#!/bin/bash
# ncu captures behind profiles/r2b_* (one GPU; summaries are made here afterwards with scripts/summarize_profiles.py).
# ncu cannot launch cooperative grids in thread-block clusters (LaunchFailed): the k-split kernel of batches <= 576 is captured
# with IKFLOW_B200_PROFILING_LAUNCH=1 (cluster launch without the cooperative attribute; profiler only).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/prof_build.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_exact.csv python bench.py --mode exact --batch 2048 --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-gpu-baseline > gpurun_out/prof_exact.log 2>&1
for b in 1024 2048 8192; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flow_inverse -s 3 -c 1 -f -o gpurun_out/r2b_flow_pp_b$b python scripts/prof_flow.py $b 5 > gpurun_out/prof_$b.log 2>&1
done
IKFLOW_B200_PROFILING_LAUNCH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:flow_inverse -s 5 -c 1 -f -o gpurun_out/r2b_flow_ks_b512 python scripts/prof_flow.py 512 8 > gpurun_out/prof_512.log 2>&1
IKFLOW_B200_PROFILING_LAUNCH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline --no-gpu-baseline > gpurun_out/prof_bench.log 2>&1
echo done
