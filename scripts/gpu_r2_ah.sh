#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/ah_build.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_flow.py > gpurun_out/ah_memcheck.log 2>&1
echo "rc $?" >> gpurun_out/ah_memcheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python scripts/sanitize_flow.py > gpurun_out/ah_synccheck.log 2>&1
echo "rc $?" >> gpurun_out/ah_synccheck.log
echo done
