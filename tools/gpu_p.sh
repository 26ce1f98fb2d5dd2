#!/bin/bash
# A/B of the variants under ray_tracing-rendering_b200/variants/ on the fused-kernel configurations, then the image tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python tools/variant_sweep.py C4env C1 C3 --reps 3
for v in "" $(ls ray_tracing-rendering_b200/variants/ | grep "^librtb200_.*\.so$" | sed 's/librtb200_//; s/\.so//'); do
  if [ -n "$v" ]; then export RTB200_LIBRARY=$PWD/ray_tracing-rendering_b200/variants/librtb200_$v.so; else unset RTB200_LIBRARY; fi
  echo "== C4 --fused variant '$v'"
  timeout 600 python tools/run_config.py C4 --fused --reps 3 | tail -n 3
  echo "== C2 --fused? C4 default variant '$v'"
  timeout 600 python tools/run_config.py C4 --reps 3 | tail -n 3
done
unset RTB200_LIBRARY
timeout 1500 python -m pytest tests/test_gpu_render.py -x -q -m gpu 2>&1 | tail -n 5
} > gpurun_out/p_sweep.log 2>&1
cat gpurun_out/p_sweep.log
