#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/q_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_flow.py -q --tb=short -k "ksplit" -x > gpurun_out/q_tests_ks.log 2>&1
for i in 1 2; do
for prec in bf16x3 fp16x3; do
IKFLOW_B200_PRECISION=$prec timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 64 512 576 >> gpurun_out/q_time.jsonl 2> /dev/null
IKFLOW_B200_PRECISION=$prec IKFLOW_B200_KSPLIT=0 timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 64 512 576 >> gpurun_out/q_time.jsonl 2> /dev/null
done
done
timeout 300 python scripts/trace_flow.py 512 6 > gpurun_out/q_trace512_ks.log 2>&1
echo done
