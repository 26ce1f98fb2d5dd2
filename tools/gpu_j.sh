#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/j_smoke.log
timeout 600 python tools/sanitize_run.py > gpurun_out/j_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/j_memcheck.log 2>&1
echo "memcheck exit $?"; tail -n 6 gpurun_out/j_memcheck.log; tail -n 3 gpurun_out/j_plain.log
