#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/ac_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/ac_tests.log 2>&1
timeout 300 python scripts/pp_check.py fetch_arm__large__mh186_9.25m 4096 2400 > gpurun_out/ac_pp_fetch.log 2>&1
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/ac_bench.json 2> gpurun_out/ac_bench.err
echo done
