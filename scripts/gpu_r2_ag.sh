#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/ag_build.log 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/ag_bench8.json 2> gpurun_out/ag_bench8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 scripts/peer_gather_check.py > gpurun_out/ag_gather8.log 2>&1
IKFLOW_B200_GATHER=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 100 --warmup 10 --no-extra --no-gpu-baseline --no-cpu-baseline > gpurun_out/ag_bench8_nccl.json 2> gpurun_out/ag_bench8_nccl.err
echo done
