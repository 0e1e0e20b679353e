#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/aa_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/aa_tests.log 2>&1
for i in 1 2; do
for prec in bf16x3 fp16x3; do
IKFLOW_B200_PRECISION=$prec timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 64 512 1024 8192 >> gpurun_out/aa_time.jsonl 2> /dev/null
done
done
echo done
