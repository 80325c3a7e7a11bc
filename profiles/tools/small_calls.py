import time, numpy as np, torch
import fimex_b200 as fb
from fimex_b200 import Method
SRC = "+proj=latlong +a=6371000 +e=0 +no_defs"
DST = "+proj=ob_tran +o_proj=longlat +lon_0=-40 +o_lat_p=22 +R=6.371e+06 +no_defs"
lon = np.arange(1440) * 0.25; lat = 90.0 - np.arange(721) * 0.25
ax = (np.arange(2000) - 999.5) * 0.0225
ci = fb.CachedInterpolation.fromProjection(Method.BILINEAR, DST, ax, ax, True, True, SRC, lon, lat, True)
ci.createReducedDomain()
for nz in (1, 4, 16):
    h_in = torch.randn((nz, ci.getInY(), ci.getInX())).pin_memory()
    h_out = torch.empty(nz * 4_000_000, dtype=torch.float32).pin_memory()
    a, b = h_in.numpy(), h_out.numpy()
    ci.interpolateValues(a, out=b)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); ci.interpolateValues(a, out=b); ts.append(time.perf_counter() - t0)
    ideal = b.nbytes / 55e9
    print(f"nz={nz:3d}: per call {np.median(ts)*1e3:7.3f} ms (min {min(ts)*1e3:.3f}), D2H alone at 55 GB/s would be {ideal*1e3:.3f} ms", flush=True)
    d_in = h_in.cuda(); d_out = torch.empty(nz * 4_000_000, device="cuda")
    ci.interpolateValues(d_in, out=d_out); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50): ci.interpolateValues(d_in, out=d_out)
    torch.cuda.synchronize()
    print(f"         device-resident call: {(time.perf_counter()-t0)/50*1e3:.3f} ms", flush=True)
