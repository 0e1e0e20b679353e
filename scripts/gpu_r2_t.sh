#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/t_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_flow.py -q --tb=short -x > gpurun_out/t_tests.log 2>&1
for i in 1 2; do
for prec in bf16x3 fp16x3; do
IKFLOW_B200_PRECISION=$prec timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 64 512 1024 8192 >> gpurun_out/t_time.jsonl 2> /dev/null
done
done
timeout 300 python scripts/trace_flow.py 512 6 > gpurun_out/t_trace512_ks.log 2>&1
timeout 600 python scripts/precision_gpu.py 256 > gpurun_out/t_precision.log 2>&1
echo done
