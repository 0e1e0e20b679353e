#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/d_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/d_tests.log 2>&1
for cs in 2 1; do
IKFLOW_B200_CLUSTER=$cs timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 512 1024 2048 4096 8192 >> gpurun_out/d_time.jsonl 2> /dev/null
done
IKFLOW_B200_RT=64 timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 2048 8192 >> gpurun_out/d_time.jsonl 2> /dev/null
IKFLOW_B200_JIT=0 timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 512 >> gpurun_out/d_time.jsonl 2> /dev/null
timeout 300 python scripts/time_flow.py fetch_arm__large__mh186_9.25m 512 4096 >> gpurun_out/d_time.jsonl 2> /dev/null
timeout 300 python scripts/time_flow.py panda__nb16__synthetic 8192 >> gpurun_out/d_time.jsonl 2> /dev/null
echo done
