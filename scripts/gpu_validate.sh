#!/bin/bash
# One GPU box: build, GPU test suite, smoke, default bench.   gpurun --timeout 2700 -- 'bash scripts/gpu_validate.sh'
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/validate_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/validate_tests.log 2>&1
timeout 120 python __graft_entry__.py --smoke > gpurun_out/validate_smoke.log 2>&1
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err
echo done
