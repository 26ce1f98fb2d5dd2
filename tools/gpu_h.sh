#!/bin/bash
# A/B of compile-time variants under ray_tracing-rendering_b200/variants/ on C5 and scene09 (+ parity of each variant)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 1200 python tools/variant_sweep.py C5 --spp 16 --reps 3
for v in "" $(ls ray_tracing-rendering_b200/variants/ | grep "^librtb200_.*\.so$" | sed 's/librtb200_//; s/\.so//'); do
  if [ -n "$v" ]; then export RTB200_LIBRARY=$PWD/ray_tracing-rendering_b200/variants/librtb200_$v.so; else unset RTB200_LIBRARY; fi
  echo "== C2 --wide variant '$v'"
  timeout 600 python tools/run_config.py C2 --spp 100 --wide --reps 2 --time --count | tail -n 2
  echo "== C5 variant '$v' counts"
  timeout 600 python tools/run_config.py C5 --spp 16 --reps 1 --time --count | tail -n 1
  if [ -n "$v" ]; then
    echo "== parity of variant '$v'"
    true
  fi
done
unset RTB200_LIBRARY
} > gpurun_out/h_sweep.log 2>&1
cat gpurun_out/h_sweep.log
