import time, os, numpy as np, torch
import fimex_b200 as fb
def timed(fn, f, reps=3):
    d = f.clone(); fn(d); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        d = f.clone(); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(d); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best, d
for shape in ((137, 202, 1440), (1, 202, 1440), (8, 2000, 2000)):
    g = torch.Generator(device="cuda").manual_seed(1)
    f = torch.randn(shape, device="cuda", generator=g) + 280
    f[torch.rand(shape, device="cuda", generator=g) < 0.05] = float("nan")
    f[..., 50:90, 100:300] = float("nan")
    for name, fn in (("fill2d(0.01,1.6,100)", lambda d: fb.fill2d_device(d, 0.01, 1.6, 100)), ("fill2d(1e-9,1.6,100)", lambda d: fb.fill2d_device(d, 1e-9, 1.6, 100)),
                     ("creepfill2d(20,2)", lambda d: fb.creepfill2d_device(d, 20, 2))):
        dt, d = timed(fn, f)
        print(f"{os.environ.get('FIMEX_B200_FILL_SIMPLE','skewed'):7s} {name:22s} {shape}: {dt*1e3:9.2f} ms, NaN left {int(torch.isnan(d).sum())}", flush=True)
