#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/f_bench.json").read().strip().splitlines()[-1])
print("C1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
for k, v in d["configs"].items():
    print(k, round(v["mpaths_per_s"], 1), "Mpaths/s", round(v["ms_per_step"], 2), "ms e2e", round(v["e2e"]["value"], 1), v["schedule"])
PY
