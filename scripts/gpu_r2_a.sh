#!/bin/bash
# round 2, GPU call A: tests, bench (all keys), precision A/B, kinematics launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/a_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -rA 2>&1 | tail -150 > gpurun_out/a_tests.log
echo "tests rc=${PIPESTATUS[0]}" >> gpurun_out/a_tests.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?" >> gpurun_out/a_bench.err
for prec in bf16x3 fp16x3; do
  for b in 512 2048 8192; do
    timeout 300 python bench.py --steps 100 --warmup 10 --batch $b --precision $prec --no-extra --no-cpu-baseline --no-gpu-baseline >> gpurun_out/a_prec_ab.jsonl 2>> gpurun_out/a_prec_ab.err
  done
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/a_smoke.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/a_launches_exact.csv python bench.py --mode exact --batch 2048 --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-gpu-baseline > gpurun_out/a_ncu_exact.log 2>&1
echo done
