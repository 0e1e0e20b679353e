#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/aj_build.log 2>&1
timeout 300 python scripts/pp_check.py panda__full__lp191_5.25m 1153 1536 2048 2304 2305 8192 > gpurun_out/aj_pp.log 2>&1
echo "rc $?" >> gpurun_out/aj_pp.log
echo done
