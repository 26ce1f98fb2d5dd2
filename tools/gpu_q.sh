#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for v in "" $(ls ray_tracing-rendering_b200/variants/ | grep "^librtb200_.*\.so$" | sed 's/librtb200_//; s/\.so//'); do
  if [ -n "$v" ]; then export RTB200_LIBRARY=$PWD/ray_tracing-rendering_b200/variants/librtb200_$v.so; else unset RTB200_LIBRARY; fi
  echo "== variant '$v'"
  timeout 600 python tools/run_config.py C5 --spp 16 --reps 3 --time | tail -n 2
  timeout 600 python tools/run_config.py C2 --spp 100 --reps 3 --time | tail -n 2
  timeout 600 python tools/run_config.py C4 --reps 3 | tail -n 2
done
} > gpurun_out/q_sweep.log 2>&1
cat gpurun_out/q_sweep.log
