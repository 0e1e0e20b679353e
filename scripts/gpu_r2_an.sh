#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/an_build.log 2>&1
IKFLOW_B200_CLUSTER_PP=2 timeout 300 python scripts/pp_check.py panda__full__lp191_5.25m 700 1024 2048 2305 8192 > gpurun_out/an_pp_cs2.log 2>&1
echo "rc $?" >> gpurun_out/an_pp_cs2.log
for i in 1 2; do
for cs in 1 2; do
IKFLOW_B200_CLUSTER_PP=$cs timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 1024 2048 4096 8192 9216 >> gpurun_out/an_time.jsonl 2> /dev/null
done
done
echo done
