import time, sys, os
import numpy as np, torch
import fimex_b200 as fb
from fimex_b200 import Method
SRC = "+proj=latlong +a=6371000 +e=0 +no_defs"
DST = "+proj=ob_tran +o_proj=longlat +lon_0=-40 +o_lat_p=22 +R=6.371e+06 +no_defs"
lon = np.arange(1440) * 0.25; lat = 90.0 - np.arange(721) * 0.25
ax = (np.arange(2000) - 999.5) * 0.0225
for name, m in (("bilinear", Method.BILINEAR), ("nearestneighbor", Method.NEAREST_NEIGHBOR), ("bicubic", Method.BICUBIC)):
    ci = fb.CachedInterpolation.fromProjection(m, DST, ax, ax, True, True, SRC, lon, lat, True)
    ci.createReducedDomain()
    cvr = fb.CachedVectorReprojection.fromProjection(fb.MIFI_VECTOR_KEEP_SIZE, SRC, DST, ax, ax, fb.LONGITUDE, fb.LATITUDE)
    nz = 137 * 4
    u = torch.randn((nz, ci.getInY(), ci.getInX()), device="cuda"); v = torch.randn_like(u)
    for rot in (cvr, None):
        ci.interpolateVector(u, v, rot); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ci.interpolateVector(u, v, rot)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        vals = 2 * nz * 4e6
        print(f"{name:16s} vector rot={rot is not None}: {ms:8.3f} ms  {vals/ms*1e-6:8.1f} Gvalues/s  {vals*4.3/ms/1e6:7.0f} GB/s", flush=True)
