#!/bin/bash
# full GPU parity suite, then the headline bench (checks that the templated fused kernels kept their speed)
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py > gpurun_out/r02b_bench_1gpu.json 2> gpurun_out/r02b_bench_1gpu.err
python bench.py --prime 2027 --no-cpu --no-regression > gpurun_out/r02_bench_p2027.json 2> gpurun_out/r02_bench_p2027.err
python - <<'PY'
import json
for f in ["r02b_bench_1gpu", "r02_bench_p2027"]:
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").readline())
        print(f, d["value"], d["e2e"]["value"], d["roofline"]["frac"], (d.get("regression") or {}).get("value"), d["roofline"]["per_kernel_ms"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r02b_bench_1gpu.err
