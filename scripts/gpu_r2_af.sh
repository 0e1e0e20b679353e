#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/af_build.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__cluster_dim_x,launch__registers_per_thread,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --replay-mode application --clock-control none --metrics $M -k regex:flow_inverse -s 5 -c 1 -f -o gpurun_out/r2b_flow_ks_b512 python scripts/prof_flow.py 512 8 > gpurun_out/af_ncu_ks512.log 2>&1
echo "rc $?" >> gpurun_out/af_ncu_ks512.log
timeout 600 ncu --replay-mode application --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/af_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline --no-gpu-baseline > gpurun_out/af_ncu_bench.log 2>&1
echo "rc $?" >> gpurun_out/af_ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flow_inverse -s 3 -c 1 -f -o gpurun_out/r2b_flow_pp_b8192 python scripts/prof_flow.py 8192 5 > gpurun_out/af_ncu_pp8192.log 2>&1
echo "rc $?" >> gpurun_out/af_ncu_pp8192.log
ls -la gpurun_out/*.ncu-rep
echo done
