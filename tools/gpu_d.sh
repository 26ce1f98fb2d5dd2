#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== sweep (default = node_min 16, no octant sort, 256x3)"
timeout 1200 python tools/variant_sweep.py C2 --spp 100 --reps 2
timeout 1200 python tools/variant_sweep.py C5 --spp 16 --reps 2
echo "== full-size steps: wide vs binary"
timeout 600 python tools/run_config.py C2 --time --reps 2
timeout 600 python tools/run_config.py C2 --time --reps 2 --binary
timeout 600 python tools/run_config.py C5 --spp 64 --time --reps 2
timeout 600 python tools/run_config.py C5 --spp 64 --time --reps 2 --binary
} > gpurun_out/d_sweep.log 2>&1
cat gpurun_out/d_sweep.log
timeout 300 python tools/run_config.py C5 --spp 8 --warm 0 > gpurun_out/d_plain_c5.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend_w -s 3 -c 1 -o gpurun_out/r02d_c5_extend -f python tools/run_config.py C5 --spp 8 --warm 0 > gpurun_out/d_ncu_c5.log 2>&1
timeout 300 python tools/run_config.py C2 --spp 40 --warm 0 > gpurun_out/d_plain_c2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend_w -s 3 -c 1 -o gpurun_out/r02d_c2_extend -f python tools/run_config.py C2 --spp 40 --warm 0 > gpurun_out/d_ncu_c2.log 2>&1
tail -n 2 gpurun_out/d_ncu_c5.log gpurun_out/d_ncu_c2.log
