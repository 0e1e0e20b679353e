#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r_tests.log 2>&1
for prec in bf16x3 fp16x3; do
IKFLOW_B200_PRECISION=$prec timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 64 512 576 >> gpurun_out/r_time.jsonl 2> /dev/null
done
timeout 600 python scripts/precision_gpu.py 256 > gpurun_out/r_precision.log 2>&1
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
echo done
