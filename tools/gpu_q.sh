#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 1200 python tools/variant_sweep.py C4env --reps 3
for v in "" $(ls ray_tracing-rendering_b200/variants/ | grep "^librtb200_.*\.so$" | sed 's/librtb200_//; s/\.so//'); do
  if [ -n "$v" ]; then export RTB200_LIBRARY=$PWD/ray_tracing-rendering_b200/variants/librtb200_$v.so; else unset RTB200_LIBRARY; fi
  echo "== C4 --fused variant '$v'"
  timeout 600 python tools/run_config.py C4 --fused --reps 3 | tail -n 2
done
} > gpurun_out/q_sweep.log 2>&1
cat gpurun_out/q_sweep.log
