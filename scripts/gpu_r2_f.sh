#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/f_build.log 2>&1
IKFLOW_B200_DEBUG=4 IKFLOW_B200_JIT=0 timeout 300 python scripts/trace_flow.py 512 6 > gpurun_out/f_trace512_nojit_d4.log 2>&1
IKFLOW_B200_DEBUG=4 timeout 300 python scripts/trace_flow.py 2048 6 > gpurun_out/f_trace2048_d4.log 2>&1
IKFLOW_B200_DEBUG=4 IKFLOW_B200_JIT=0 IKFLOW_B200_CLUSTER=1 timeout 300 python scripts/trace_flow.py 512 6 > gpurun_out/f_trace512_nojit_cs1_d4.log 2>&1
echo done
