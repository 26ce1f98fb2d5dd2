#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 1200 python tools/variant_sweep.py C5 --spp 16 --reps 3
for v in "" psort topc both; do
  if [ -n "$v" ]; then export RTB200_LIBRARY=$PWD/ray_tracing-rendering_b200/variants/librtb200_$v.so; else unset RTB200_LIBRARY; fi
  echo "== C2 --wide variant '$v'"
  timeout 600 python tools/run_config.py C2 --spp 100 --wide --reps 2 --time --count | tail -n 2
  echo "== C5 variant '$v' counts"
  timeout 600 python tools/run_config.py C5 --spp 16 --reps 1 --time --count | tail -n 1
done
unset RTB200_LIBRARY
echo "== upload phases"
RTB200_TIMING=1 timeout 600 python tools/run_config.py C4env --spp 2 --reps 1 2>&1 | grep "rtb200" | tail -n 30
RTB200_TIMING=1 timeout 600 python tools/run_config.py C5 --spp 1 --reps 1 2>&1 | grep "rtb200" | tail -n 30
} > gpurun_out/h_sweep.log 2>&1
cat gpurun_out/h_sweep.log
