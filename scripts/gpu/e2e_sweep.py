"""End-to-end (host buffers) throughput of fhesi_mult_relin_host under different pipeline chunk schedules
(FHESI_PIPE_SCHED, read at every call), batch 8192, logQ = 256, p = 1019.  One line per schedule."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "fhe-si_b200"))
import pyfhesi  # noqa: E402

SCHEDS = ["", "*384", "*768", "*1024", "*1536", "*2048", "*4096", "192,*768,192", "384,*1024", "*576", "*8192"]


def main():
    B, logq, p = 8192, 256, 1019
    dev = pyfhesi.Context(p - 1, logq, p, 3, 1, 0)
    rng = np.random.default_rng(7)
    n, W, D = dev.n, dev.W, dev.D
    ksw = dev.ksw_create(rng.integers(0, 2**32, size=(3 * D, n, W), dtype=np.uint32),
                         rng.integers(0, 2**32, size=(3 * D, n, W), dtype=np.uint32), 3)
    ha = dev.host_alloc((B, 2, n, W), np.uint32, write_combined=True)
    hb = dev.host_alloc((B, 2, n, W), np.uint32, write_combined=True)
    ho = dev.host_alloc((B, 2, n, W), np.uint32, write_combined=False)
    ha[...] = rng.integers(0, 2**32, size=ha.shape, dtype=np.uint32)
    hb[...] = rng.integers(0, 2**32, size=hb.shape, dtype=np.uint32)

    ho2 = dev.host_alloc((B, 2, n, W), np.uint32, write_combined=False)
    ASYNC = "--blocking" not in sys.argv

    def call(i=0):
        if ASYNC:
            dev._ck(dev.lib.fhesi_mult_relin_host_async(dev.h, ksw, ha.ctypes.data, hb.ctypes.data,
                                                        (ho, ho2)[i & 1].ctypes.data, B))
        else:
            dev._ck(dev.lib.fhesi_mult_relin_host(dev.h, ksw, ha.ctypes.data, hb.ctypes.data, ho.ctypes.data, B))

    for s in SCHEDS + [""]:
        if s:
            os.environ["FHESI_PIPE_SCHED"] = s
        else:
            os.environ.pop("FHESI_PIPE_SCHED", None)
        for i in range(2):
            call(i)
        dev.sync_all()
        t0 = time.perf_counter()
        reps = 8
        for i in range(reps):
            call(i)
        dev.sync_all()
        ms = (time.perf_counter() - t0) * 1e3 / reps
        print(f"sched {s or 'default':40s} {ms:7.3f} ms/step  {B / ms * 1e3:9.0f} ops/s", flush=True)


if __name__ == "__main__":
    main()
