#!/bin/bash
# final-code measurements of round 2: GPU suite, default bench, launch list and full captures of the
# dominant kernels (each ncu run only after the same command exited 0 without ncu)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$1" != "nopytest" ]; then
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/g_pytest.log
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/g_pytest.log | tail -n 20
fi
mkdir -p /tmp/ncu
timeout 900 python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench exit $?"
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/g_bench_short.json 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/g_ncu_launch.log 2>&1
# the reports stay on the box (gpurun_out is capped at 64 MiB): their raw and source pages come back as CSV
for cfg in "C1 400 k_fused c1_fused 0" "C5 8 k_extend_w c5_extend_w 3" "C5 8 k_connect_w c5_connect_w 3" "C2 40 k_extend c2_extend 3"; do
    set -- $cfg
    timeout 300 python tools/run_config.py $1 --spp $2 --warm 0 > gpurun_out/g_plain_$4.log 2>&1 && \
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$3\$" -s $5 -c 1 -o /tmp/ncu/r02_$4 -f \
        python tools/run_config.py $1 --spp $2 --warm 0 > gpurun_out/g_ncu_$4.log 2>&1
    tail -n 1 gpurun_out/g_ncu_$4.log
    ncu -i /tmp/ncu/r02_$4.ncu-rep --page raw --csv > gpurun_out/r02_$4.raw.csv 2>/dev/null
    ncu -i /tmp/ncu/r02_$4.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02_$4.source.csv 2>/dev/null
done
du -sh gpurun_out
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/g_bench.json").read().strip().splitlines()[-1])
    print("C1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
    for k, v in d["configs"].items():
        print(k, round(v["mpaths_per_s"], 1), "Mpaths/s", round(v["ms_per_step"], 2), "ms e2e", round(v["e2e"]["value"], 1), v["schedule"])
except Exception as e:
    print("parse failed", e)
PY
