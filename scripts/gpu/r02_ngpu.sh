#!/bin/bash
# bench.py on N GPUs of one box under torchrun, as the driver launches it; $1 = N, $2 = tag
N=$1; TAG=${2:-r02g}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
python - <<PY
import json
ls = open("gpurun_out/${TAG}_bench_${N}gpu.json").read().splitlines()
d = json.loads(ls[-1])
print(len(ls), "stdout line(s);", d["n_gpus"], "GPUs:", round(d["value"]), "ops/s, e2e", round(d["e2e"]["value"]), "bound",
      round(d["e2e"]["host_copy_bound"]["ops_s"]), "regression", d["regression"]["value"], "exchange us", (d["exchange"] or {}).get("us"))
PY
