#!/bin/bash
TOOL=${1:-memcheck}
export SPZB200_NO_REBUILD=1
mkdir -p gpurun_out
python scripts/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 --log-file gpurun_out/sanitizer_${TOOL}.log python scripts/sanitize_case.py > gpurun_out/sanitize_${TOOL}_stdout.log 2>&1
echo "sanitizer $TOOL rc=$?"
tail -3 gpurun_out/sanitize_plain.log; tail -8 gpurun_out/sanitizer_${TOOL}.log; tail -2 gpurun_out/sanitize_${TOOL}_stdout.log
