#!/bin/bash
# 2 GPUs: bench.py under torchrun (fused gather) and with the NCCL fallback
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/j_build.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/j_bench2.json 2> gpurun_out/j_bench2.err
echo "rc=$?" >> gpurun_out/j_bench2.err
IKFLOW_B200_GATHER=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 10 --no-extra > gpurun_out/j_bench2_nccl.json 2> gpurun_out/j_bench2_nccl.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/j_ref2.json 2>/dev/null
echo done
