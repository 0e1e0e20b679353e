#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/ai_build.log 2>&1
timeout 900 python scripts/precision_gpu.py 2400 > gpurun_out/ai_precision_pp.log 2>&1
IKFLOW_B200_PP=0 timeout 900 python scripts/precision_gpu.py 2400 > gpurun_out/ai_precision_nopp.log 2>&1
echo done
