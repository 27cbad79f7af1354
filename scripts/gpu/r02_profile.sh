#!/bin/bash
# ncu evidence for the hot path as built now: (1) launch list of one bench step, (2) one --set full capture of each of the
# five hot kernels (first launch after the warm-up steps of scripts/gpu/ks_variants.py --single), (3) an un-profiled bench
TAG=${1:-r02c}
# (the filter keeps the five kernels of the device-resident step and drops the set-up's encryptions; the first 15
# matching launches are the warm-up step and the two timed steps, before the end-to-end leg's chunked launches)
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:k_fused|k_crt_direct|k_crt_split|k_residues' \
    --launch-count 15 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-regression > gpurun_out/${TAG}_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_fused|k_crt|k_residues' \
    --launch-skip 15 --launch-count 5 -f -o gpurun_out/${TAG}_full python scripts/gpu/ks_variants.py --single \
    > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log
python bench.py > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_1gpu.json').readline())
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['regression']['value'], d['roofline']['per_kernel_ms'])"
