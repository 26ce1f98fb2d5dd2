#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/e_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/e_pytest.log
tail -n 6 gpurun_out/e_pytest.log
timeout 900 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
echo "bench exit $?"
tail -c 1500 gpurun_out/e_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/e_bench.json").read().strip().splitlines()[-1])
    print("C1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["bound"], d["roofline"]["frac"])
    for k, v in d["configs"].items():
        print(k, round(v["mpaths_per_s"], 1), "Mpaths/s", round(v["ms_per_step"], 2), "ms e2e", round(v["e2e"]["value"], 1), v["schedule"], v["roofline"]["bound"], round(v["roofline"]["frac"], 3), "spp", v["spp_run"])
    for k, v in d["strong"].items():
        print("strong", k, round(v["mpaths_per_s"], 1), round(v["ms_per_step"], 2))
    print("cpu", d["cpu_baseline"])
except Exception as e:
    print("parse failed", e)
PY
