#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 1200 python tools/variant_sweep.py C3 C1 --reps 3
} > gpurun_out/q_sweep.log 2>&1
cat gpurun_out/q_sweep.log
