#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/ae_build.log 2>&1
TRACE_PRECISION=bf16x3 timeout 300 python scripts/trace_flow.py 8192 4 > gpurun_out/ae_trace8192_pp.log 2>&1
echo done
