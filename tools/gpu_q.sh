#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 600 python tools/run_config.py C4env --reps 2 --time | tail -n 2
timeout 600 python tools/run_config.py C4env --wavefront --reps 2 --time | tail -n 2
timeout 600 python tools/run_config.py C4 --reps 2 --time | tail -n 2
timeout 600 python tools/run_config.py C4 --fused --reps 2 --time | tail -n 2
} > gpurun_out/q_sweep.log 2>&1
cat gpurun_out/q_sweep.log
