import torch, time
n = 532676608
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n//2, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device='cuda'); d_out = torch.empty(n//2, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t) / 5
run(1, 1)
a, b, c = run(1, 0), run(0, 1), run(1, 1)
print(f"h2d alone {n/a/1e9:.1f} GB/s, d2h alone {n/2/b/1e9:.1f} GB/s, both: {c*1e3:.2f} ms -> h2d {n/c/1e9:.1f} GB/s + d2h {n/2/c/1e9:.1f} GB/s")
