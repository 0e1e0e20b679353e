#!/bin/bash
# N GPUs of one box (N = $1, default 2): fused-gather check and the bench under torchrun.
#   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/gpu_multi.sh 2'
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/multi_build.log 2>&1
[ "$N" = 2 ] && timeout 600 python -m pytest tests/test_gpu_multi.py -q --tb=short > gpurun_out/multi_tests.log 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/multi_bench$N.json 2> gpurun_out/multi_bench$N.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 scripts/peer_gather_check.py > gpurun_out/multi_gather$N.log 2>&1
IKFLOW_B200_GATHER=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 100 --warmup 10 --no-extra --no-gpu-baseline --no-cpu-baseline > gpurun_out/multi_bench${N}_nccl.json 2> gpurun_out/multi_bench${N}_nccl.err
echo done
