"""Time fhesi_mult_relin_dev (batch 8192, logQ = 256, p = 1019) with several builds of the CUDA library that differ in
compile-time switches of the fused kernels (scripts/gpu/build_variants.sh).  Each variant is also checked against the
default build byte for byte.  Prints one line per variant: ops/s and per-kernel ms (CUDA events on the library's stream)."""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "fhe-si_b200"))
import pyfhesi  # noqa: E402


def run(lib, B=8192, steps=6, logq=256, p=1019, ref=None):
    import torch
    dev = pyfhesi.Context(p - 1, logq, p, 3, 1, 0, lib_path=lib)
    rng = np.random.default_rng(7)
    n, W, D = dev.n, dev.W, dev.D
    kb = rng.integers(0, 2**32, size=(3 * D, n, W), dtype=np.uint32)
    kA = rng.integers(0, 2**32, size=(3 * D, n, W), dtype=np.uint32)
    ksw = dev.ksw_create(kb, kA, 3)
    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.randint(-2**31, 2**31 - 1, (B, 2, n, W), device="cuda", dtype=torch.int32, generator=g)  # full-range coefficients
    b = torch.randint(-2**31, 2**31 - 1, (B, 2, n, W), device="cuda", dtype=torch.int32, generator=g)
    out = torch.empty_like(a)
    torch.cuda.synchronize()
    for _ in range(3):
        dev.mult_relin_dev(ksw, a, b, out, B)
    dev.sync()
    dev.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.Stream()
    dev.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            dev.mult_relin_dev(ksw, a, b, out, B)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    prof = {k: round(v[1] / v[0], 4) for k, v in dev.profile_report().items()}
    res = out[:64].cpu().numpy()
    ok = None if ref is None else bool(np.array_equal(res, ref))
    dev.set_stream(0)
    dev.close()
    return {"lib": os.path.basename(lib), "ops_s": round(B / ms * 1e3), "ms_step": round(ms, 4), "same_as_default": ok,
            "per_launch_ms": prof}, res


def main():
    default = os.path.join(ROOT, "fhe-si_b200", "libfhesi_b200.so")
    prime = 2027 if "--p2027" in sys.argv else 1019
    global run
    run0 = run
    run = lambda lib, **kw: run0(lib, p=prime, **kw)
    if "--single" in sys.argv:  # the default build only (what ncu is pointed at: scripts/gpu/r02_profile.sh)
        print(json.dumps(run(default, steps=2)[0]), flush=True)
        return
    r, ref = run(default)
    print(json.dumps(r), flush=True)
    for lib in sorted(glob.glob(os.path.join(ROOT, "scripts", "gpu", "variants", "*.so"))):
        try:
            r, _ = run(lib, ref=ref)
        except Exception as e:  # a variant that fails to launch must not hide the others
            r = {"lib": os.path.basename(lib), "error": str(e)}
        print(json.dumps(r), flush=True)
    r, _ = run(default, ref=ref)
    print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
