#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/n_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/n_tests.log 2>&1
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
timeout 300 python __graft_entry__.py --smoke > gpurun_out/n_smoke.log 2>&1
echo done
