#!/bin/bash
# 2 GPUs: fused gather vs NCCL all-gather
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/i_build.log 2>&1
nvidia-smi topo -m > gpurun_out/i_topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/peer_gather_check.py > gpurun_out/i_peer.log 2>&1
echo "rc=$?" >> gpurun_out/i_peer.log
echo done
