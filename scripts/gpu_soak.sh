#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/soak_build.log 2>&1
for b in 64 512 576 1024 2048 2400 8192; do
n=3000; [ $b -ge 2400 ] && n=1000
timeout 300 python scripts/stress_flow.py $b $n 2>/dev/null | tail -n 1 >> gpurun_out/soak.log
IKFLOW_B200_PRECISION=bf16x3 timeout 300 python scripts/stress_flow.py $b $n 2>/dev/null | tail -n 1 >> gpurun_out/soak.log
done
echo done
