#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_flow.py tests/test_gpu_reference_fixtures.py -q -s --tb=long -k "stress or config3" > gpurun_out/c_tests.log 2>&1
echo done
