#!/bin/bash
# round 2, GPU call B: cluster multicast correctness + A/B timing, precision study, full test suite
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/b_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_flow.py -q -k "multicast" -rA 2>&1 | tail -40 > gpurun_out/b_cluster_tests.log
for cs in 1 2 4 1 2 4; do
  IKFLOW_B200_CLUSTER=$cs timeout 300 python scripts/time_flow.py panda__full__lp191_5.25m 512 1024 2048 8192 >> gpurun_out/b_cluster_ab.jsonl 2>> gpurun_out/b_cluster_ab.err
done
for cs in 1 2 4; do
  IKFLOW_B200_CLUSTER=$cs timeout 300 python scripts/time_flow.py fetch_arm__large__mh186_9.25m 512 4096 >> gpurun_out/b_cluster_ab.jsonl 2>> gpurun_out/b_cluster_ab.err
done
timeout 600 python scripts/precision_gpu.py 256 > gpurun_out/b_precision.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -rA 2>&1 | tail -60 > gpurun_out/b_tests.log
echo done
