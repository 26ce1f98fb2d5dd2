#!/bin/bash
# grouped boxes (RTB_OPT_GROUP_BOXES) on scene09: binary and 4-wide kernels, grouped and not; then the GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for mode in "" "--wide" "--binary"; do
for grp in "" "--nogroup"; do
  echo "== C2 $mode $grp"
  timeout 600 python tools/run_config.py C2 --spp 100 $mode $grp --reps 2 --time --count | tail -n 2
done
done
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 6
} > gpurun_out/q_sweep.log 2>&1
cat gpurun_out/q_sweep.log
