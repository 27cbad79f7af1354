#!/bin/bash
# GPU parity tests, then the p = 2027 family (N = 2048) through bench.py, fused and generic kernels
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --prime 2027 --no-cpu --no-regression > gpurun_out/r02_bench_p2027.json 2> gpurun_out/r02_bench_p2027.err
FHESI_NO_FUSED=1 python bench.py --prime 2027 --no-cpu --no-regression --steps 3 --warmup 3 > gpurun_out/r02_bench_p2027_generic.json 2> gpurun_out/r02_bench_p2027_generic.err
python - <<'PY'
import json
for f in ["r02_bench_p2027", "r02_bench_p2027_generic"]:
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").readline())
        print(f, d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["per_kernel_ms"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r02_bench_p2027.err
