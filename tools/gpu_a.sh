#!/bin/bash
# first GPU pass of round 2: sanity of the wide traversal, parity suite, variant sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
{
echo "== sanity wide vs binary"
timeout 300 python tools/run_config.py C2 --spp 20 --time --count
timeout 300 python tools/run_config.py C2 --spp 20 --time --count --binary
timeout 300 python tools/run_config.py C5 --spp 4 --time --count
timeout 300 python tools/run_config.py C5 --spp 4 --time --count --binary
timeout 300 python tools/run_config.py C4 --time --wavefront --reps 2
timeout 300 python tools/run_config.py C4 --time --wavefront --reps 2 --binary
} > gpurun_out/a_sanity.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest.log
{
timeout 1200 python tools/variant_sweep.py C2 --spp 100 --reps 2
timeout 1200 python tools/variant_sweep.py C5 --spp 8 --reps 2
} > gpurun_out/a_sweep.log 2>&1
tail -3 gpurun_out/a_pytest.log
cat gpurun_out/a_sanity.log
cat gpurun_out/a_sweep.log
