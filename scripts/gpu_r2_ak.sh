#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/ak_build.log 2>&1
timeout 300 python scripts/pp_check.py panda__full__lp191_5.25m 577 800 1024 1152 1153 2048 > gpurun_out/ak_pp.log 2>&1
echo "rc $?" >> gpurun_out/ak_pp.log
echo done
