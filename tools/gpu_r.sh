#!/bin/bash
# one full ncu capture of k_extend_w on scene09 (the 4-wide kernels forced), exported as CSV pages
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
timeout 300 python tools/run_config.py C2 --spp 40 --warm 0 --wide > gpurun_out/r_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_extend_w\$" -s 3 -c 1 -o /tmp/ncu/r_cap -f \
    python tools/run_config.py C2 --spp 40 --warm 0 --wide > gpurun_out/r_ncu.log 2>&1
tail -n 2 gpurun_out/r_ncu.log
ncu -i /tmp/ncu/r_cap.ncu-rep --page raw --csv > gpurun_out/r_c2_extend_w.raw.csv 2>/dev/null
ncu -i /tmp/ncu/r_cap.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r_c2_extend_w.source.csv 2>/dev/null
ls -la gpurun_out/r_c2_extend_w.*
