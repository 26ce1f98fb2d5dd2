#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for v in "" $(ls ray_tracing-rendering_b200/variants/ | grep "^librtb200_.*\.so$" | sed 's/librtb200_//; s/\.so//'); do
  if [ -n "$v" ]; then export RTB200_LIBRARY=$PWD/ray_tracing-rendering_b200/variants/librtb200_$v.so; else unset RTB200_LIBRARY; fi
  echo "== variant '$v'"
  timeout 900 python tools/fused_catalogue_probe.py 2>&1 | tail -n 80
done
} > gpurun_out/q_sweep.log 2>&1
cat gpurun_out/q_sweep.log
