#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/y_build.log 2>&1
TRACE_PAIRS=1 timeout 300 python scripts/trace_flow.py 512 6 > gpurun_out/y_trace512_ks.log 2>&1
timeout 300 python scripts/trace_flow.py 8192 4 > gpurun_out/y_trace8192.log 2>&1
echo done
