#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/ab_build.log 2>&1
timeout 240 python scripts/pp_check.py panda__full__lp191_5.25m 2305 4096 8192 9216 > gpurun_out/ab_pp.log 2>&1
echo "rc $?" >> gpurun_out/ab_pp.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv >> gpurun_out/ab_pp.log 2>&1
echo done
