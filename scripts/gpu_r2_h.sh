#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/h_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/h_tests.log 2>&1
timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 512 1024 2048 8192 >> gpurun_out/h_time.jsonl 2> /dev/null
IKFLOW_B200_JIT=0 timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 512 >> gpurun_out/h_time.jsonl 2> /dev/null
echo done
