#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/j_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_flow.py tests/test_gpu_solver.py tests/test_gpu_reference_fixtures.py -q --tb=short -x -k "tail_split or pingpong or exact or config3 or 8192" > gpurun_out/j_tests.log 2>&1
for i in 1 2; do
for ts in 1 0; do
IKFLOW_B200_TAIL_SPLIT=$ts timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 5000 6144 8192 20440 >> gpurun_out/j_time.jsonl 2> /dev/null
done
done
timeout 300 python bench.py --mode exact --batch 2048 --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-gpu-baseline > gpurun_out/j_exact.json 2>/dev/null
IKFLOW_B200_TAIL_SPLIT=0 timeout 300 python bench.py --mode exact --batch 2048 --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-gpu-baseline > gpurun_out/j_exact0.json 2>/dev/null
echo done
