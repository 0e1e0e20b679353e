#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/z_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q --tb=short > gpurun_out/z_tests_multi.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/z_bench2.json 2> gpurun_out/z_bench2.err
IKFLOW_B200_GATHER=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 --no-extra --no-gpu-baseline --no-cpu-baseline > gpurun_out/z_bench2_nccl.json 2> gpurun_out/z_bench2_nccl.err
echo done
