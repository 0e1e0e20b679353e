#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/ad_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_flow.py -q --tb=short -k "pingpong or ksplit or last_kernel" > gpurun_out/ad_tests.log 2>&1
for i in 1 2; do
IKFLOW_B200_PRECISION=fp16x3 timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 16 64 128 512 >> gpurun_out/ad_time.jsonl 2> /dev/null
IKFLOW_B200_PRECISION=fp16x3 IKFLOW_B200_KSPLIT=0 timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 16 64 128 512 >> gpurun_out/ad_time.jsonl 2> /dev/null
done
echo done
