#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s_build.log 2>&1
timeout 300 python scripts/trace_flow.py 8192 4 > gpurun_out/s_trace8192.log 2>&1
timeout 300 python scripts/trace_flow.py 2048 4 > gpurun_out/s_trace2048.log 2>&1
timeout 300 python scripts/trace_flow.py 1184 4 > gpurun_out/s_trace1184.log 2>&1
for prec in bf16x3 fp16x3; do
IKFLOW_B200_PRECISION=$prec timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 1024 1184 2048 2304 4096 8192 9216 16384 >> gpurun_out/s_time.jsonl 2> /dev/null
done
echo done
