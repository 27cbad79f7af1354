"""Host-copy probe: H2D from ordinary pinned memory against write-combined pinned memory (cudaHostAllocWriteCombined),
alone and with the D2H direction running, on every visible GPU concurrently when launched under torchrun.
One line per rank."""
import ctypes
import os
import time

import torch

rank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(rank)
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
n = 532676608


def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    return p.value


d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
h_out = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
bufs = {"pinned": host_alloc(n, 0), "write_combined": host_alloc(n, 4)}
for p in bufs.values():
    ctypes.memset(p, 1, n)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]


def run(src, h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            rt.cudaMemcpyAsync(d_in.data_ptr(), src, n, 1, s1.cuda_stream)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
    barrier = dist.barrier
else:
    barrier = lambda: None
out = {}
for name, p in bufs.items():
    run(p, 1, 1, 2)
    barrier()
    a = run(p, 1, 0)
    barrier()
    c = run(p, 1, 1)
    out[name] = f"h2d alone {n / a / 1e9:.1f} GB/s; both directions {c * 1e3:.2f} ms (h2d {n / c / 1e9:.1f} + d2h {n / 2 / c / 1e9:.1f} GB/s)"
print(f"rank {rank}: {out}", flush=True)
