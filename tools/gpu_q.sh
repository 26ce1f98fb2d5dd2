#!/bin/bash
# A/B of traversal variants on C5 (+ the trace-batch parity tests of the default build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 1200 python tools/variant_sweep.py C5 --spp 16 --reps 3
timeout 600 python tools/run_config.py C5 --spp 16 --reps 1 --time --count | tail -n 1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_render.py -x -q -m gpu 2>&1 | tail -n 5
} > gpurun_out/q_sweep.log 2>&1
cat gpurun_out/q_sweep.log
