import torch, time
n = 2_192_000_000 // 4
d = torch.empty(n, device="cuda", dtype=torch.float32)
h = torch.empty(n, dtype=torch.float32, pin_memory=True)
for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"{name} 2.19 GB pinned: {dt*1e3:.1f} ms  {n*4/dt/1e9:.1f} GB/s", flush=True)
# two streams concurrently, halves
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h[:n//2].copy_(d[:n//2], non_blocking=True)
    with torch.cuda.stream(s2): h[n//2:].copy_(d[n//2:], non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"D2H two streams: {n*4/dt/1e9:.1f} GB/s")
