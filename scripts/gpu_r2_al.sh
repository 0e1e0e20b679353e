#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/al_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/al_tests.log 2>&1
timeout 120 python __graft_entry__.py --smoke > gpurun_out/al_smoke.log 2>&1
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/al_bench.json 2> gpurun_out/al_bench.err
echo done
