#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== wide vs binary"
timeout 300 python tools/run_config.py C2 --spp 100 --time --reps 2
timeout 300 python tools/run_config.py C2 --spp 100 --time --reps 2 --binary
timeout 300 python tools/run_config.py C5 --spp 16 --time --reps 2
timeout 300 python tools/run_config.py C5 --spp 16 --time --reps 2 --binary
} > gpurun_out/b_sanity.log 2>&1
cat gpurun_out/b_sanity.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/b_pytest.log
tail -15 gpurun_out/b_pytest.log
timeout 300 python tools/run_config.py C5 --spp 8 --warm 0 > gpurun_out/b_plain_c5.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend_w -s 3 -c 1 -o gpurun_out/r02b_c5_extend -f python tools/run_config.py C5 --spp 8 --warm 0 > gpurun_out/b_ncu_c5.log 2>&1
timeout 300 python tools/run_config.py C2 --spp 40 --warm 0 > gpurun_out/b_plain_c2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend_w -s 3 -c 1 -o gpurun_out/r02b_c2_extend -f python tools/run_config.py C2 --spp 40 --warm 0 > gpurun_out/b_ncu_c2.log 2>&1
tail -3 gpurun_out/b_ncu_c5.log gpurun_out/b_ncu_c2.log
