#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== binary reference"
timeout 300 python tools/run_config.py C2 --spp 100 --time --reps 2 --binary
timeout 300 python tools/run_config.py C5 --spp 16 --time --reps 2 --binary
echo "== wide default"
timeout 300 python tools/run_config.py C2 --spp 100 --time --reps 2
timeout 300 python tools/run_config.py C5 --spp 16 --time --reps 2
echo "== sweep"
timeout 1200 python tools/variant_sweep.py C2 --spp 100 --reps 2
timeout 1200 python tools/variant_sweep.py C5 --spp 16 --reps 2
} > gpurun_out/c_sweep.log 2>&1
cat gpurun_out/c_sweep.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest.log
tail -n 8 gpurun_out/c_pytest.log
