#!/bin/bash
# A/B of the variants under ray_tracing-rendering_b200/variants/ on the fused-kernel configurations, then the image tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python tools/variant_sweep.py C1 C3 C4env --reps 3
timeout 1500 python -m pytest tests/test_gpu_render.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -n 5
} > gpurun_out/p_sweep.log 2>&1
cat gpurun_out/p_sweep.log
