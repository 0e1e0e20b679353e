#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/g_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/g_tests.log 2>&1
for i in 1 2; do
timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 512 1024 2048 8192 >> gpurun_out/g_time.jsonl 2> /dev/null
IKFLOW_B200_JIT=0 timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 512 >> gpurun_out/g_time.jsonl 2> /dev/null
IKFLOW_B200_JIT=0 IKFLOW_B200_CLUSTER=1 timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 512 >> gpurun_out/g_time.jsonl 2> /dev/null
done
timeout 300 python scripts/time_flow.py fetch_arm__large__mh186_9.25m 512 4096 >> gpurun_out/g_time.jsonl 2> /dev/null
timeout 300 python scripts/time_flow.py panda__nb16__synthetic 8192 >> gpurun_out/g_time.jsonl 2> /dev/null
IKFLOW_B200_JIT=0 timeout 300 python scripts/trace_flow.py 512 6 > gpurun_out/g_trace512_nojit.log 2>&1
timeout 300 python scripts/trace_flow.py 512 6 > gpurun_out/g_trace512.log 2>&1
timeout 300 python scripts/trace_flow.py 2048 6 > gpurun_out/g_trace2048.log 2>&1
echo done
