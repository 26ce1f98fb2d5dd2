#!/bin/bash
# N-GPU pass: multi-GPU tests (library group API, C++ renderer on N GPUs) + bench at N ranks
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m${N}_gpus.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_distributed_cpu.py -q > gpurun_out/m${N}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/m${N}_pytest.log
tail -n 5 gpurun_out/m${N}_pytest.log
NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
echo "bench exit $?"
tail -c 600 gpurun_out/m${N}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/m${N}_bench.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "C1", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1))
    for k, v in d.get("configs", {}).items():
        print(k, round(v["mpaths_per_s"], 1), "Mpaths/s", round(v["ms_per_step"], 2), "ms")
    for k, v in d.get("strong", {}).items():
        print("strong", k, round(v["mpaths_per_s"], 1), "Mpaths/s", round(v["ms_per_step"], 2), "ms spp_run", v["spp_run"])
except Exception as e:
    print("parse failed", e)
PY
