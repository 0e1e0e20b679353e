#!/bin/bash
# 8 GPUs: fused gather check + bench (fused, nccl)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/p_build.log 2>&1
N=${NGPU:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/peer_gather_check.py > gpurun_out/p_peer$N.log 2>&1
echo "rc=$?" >> gpurun_out/p_peer$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/p_bench$N.json 2> gpurun_out/p_bench$N.err
echo "rc=$?" >> gpurun_out/p_bench$N.err


echo done
