#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/u_build.log 2>&1
TRACE_PAIRS=1 timeout 300 python scripts/trace_flow.py 512 6 > gpurun_out/u_trace512_ks.log 2>&1
echo done
