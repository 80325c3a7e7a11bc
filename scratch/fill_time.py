import time, numpy as np, torch
import fimex_b200 as fb
rng = np.random.default_rng(3)
for shape in ((137, 202, 1440), (8, 2000, 2000)):
    ny, nx = shape[-2:]
    f = torch.randn(shape, device="cuda") + 280
    f[torch.rand(shape, device="cuda") < 0.05] = float("nan")
    f[..., 50:90, 100:300] = float("nan")
    for name, fn in (("fill2d(0.01,1.6,100)", lambda d: fb.fill2d_device(d, 0.01, 1.6, 100)), ("creepfill2d(20,2)", lambda d: fb.creepfill2d_device(d, 20, 2))):
        d = f.clone(); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(d); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{name:22s} {shape}: {dt*1e3:9.1f} ms, NaN left {int(torch.isnan(d).sum())}", flush=True)
