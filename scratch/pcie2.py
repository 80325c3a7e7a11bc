import torch, time
n = 2_192_000_000 // 4; m = 160_000_000 // 4
d = torch.empty(n, device="cuda"); h = torch.empty(n).pin_memory()
d2 = torch.empty(m, device="cuda"); h2 = torch.empty(m).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / 5
run(True)
print("D2H 2.19 GB alone: %.2f ms" % (run(False) * 1e3))
print("D2H 2.19 GB + concurrent H2D 0.16 GB: %.2f ms" % (run(True) * 1e3))
# chunked D2H: 13 copies of 176 MB back to back in one stream
c = n // 13
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    for k in range(13):
        with torch.cuda.stream(s1): h[k*c:(k+1)*c].copy_(d[k*c:(k+1)*c], non_blocking=True)
torch.cuda.synchronize(); print("D2H 13 x 169 MB in one stream: %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
