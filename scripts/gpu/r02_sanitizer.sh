#!/bin/bash
# compute-sanitizer on the kernels that are new in round 2: general-m remainder rows and automorphism matrices, the fused
# N = 2048 kernels, the windowed CRT with its exact fallback, cp.async digit staging, single-group CTAs
S="compute-sanitizer --error-exitcode 7"
run() { echo "== $1 :: $2"; timeout 1500 $S --tool $1 python -m pytest tests/test_gpu_parity.py -x -q -k "$2" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard" | head -8; }
run memcheck "general_m and (m36 or m105 or m1320)"
run memcheck "crt_direct_paths and (cfg1 or cfg3)"
run memcheck "fused_2048"
run memcheck "general_m and m1285"
run racecheck "test_mult_relin and cfg3"
run racecheck "general_m and m1320"
