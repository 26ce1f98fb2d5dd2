#!/bin/bash
# one full ncu capture of k_fused on a configuration ($1, spp $2), exported as CSV pages
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
CFG=${1:-C4env}; SPP=${2:-50}; K=${3:-k_fused}; SKIP=${4:-0}
timeout 300 python tools/run_config.py $CFG --spp $SPP --warm 0 > gpurun_out/n_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -s $SKIP -c 1 -o /tmp/ncu/n_cap -f \
    python tools/run_config.py $CFG --spp $SPP --warm 0 > gpurun_out/n_ncu.log 2>&1
tail -n 2 gpurun_out/n_ncu.log
ncu -i /tmp/ncu/n_cap.ncu-rep --page raw --csv > gpurun_out/n_${CFG}_${K}.raw.csv 2>/dev/null
ncu -i /tmp/ncu/n_cap.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/n_${CFG}_${K}.source.csv 2>/dev/null
ls -la gpurun_out/n_${CFG}_${K}.*
