#!/bin/bash
# 8 GPUs of one box: the host-copy probe (ordinary vs write-combined pinned memory) on all ranks at once, then the bench
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR scripts/pcie_probe_wc.py 2>/dev/null | sort | tee gpurun_out/r02c_pcie_probe_8gpu.txt
$TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02c_bench_8gpu.json 2> gpurun_out/r02c_bench_8gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/r02c_bench_8gpu.json').readline())
print(d['value'], d['e2e']['value'], d['e2e']['host_copy_bound'], d['regression']['value'], d['exchange'])"
