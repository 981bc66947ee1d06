#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sanitize_target.py > gpurun_out/r02v_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_target.py > gpurun_out/r02v_memcheck.log 2>&1
echo "rc=$?" >> gpurun_out/r02v_memcheck.log
tail -3 gpurun_out/r02v_plain.log; tail -8 gpurun_out/r02v_memcheck.log
