#!/bin/bash
mkdir -p gpurun_out
./scripts/ubench/dsmem_reduce > gpurun_out/l_dsmem_reduce.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/l_build.log 2>&1
IKFLOW_B200_CLUSTER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:flow_inverse -s 3 -c 1 -f -o gpurun_out/r2_flow_b8192 python scripts/prof_flow.py 8192 5 > gpurun_out/l_ncu_8192.log 2>&1
echo done
