import time, numpy as np, torch, ctypes as C
import fimex_b200 as fb
from fimex_b200 import Method
SRC = "+proj=latlong +a=6371000 +e=0 +no_defs"
DST = "+proj=ob_tran +o_proj=longlat +lon_0=-40 +o_lat_p=22 +R=6.371e+06 +no_defs"
lon = np.arange(1440) * 0.25; lat = 90.0 - np.arange(721) * 0.25
ax = (np.arange(2000) - 999.5) * 0.0225
ci = fb.CachedInterpolation.fromProjection(Method.BILINEAR, DST, ax, ax, True, True, SRC, lon, lat, True)
ci.createReducedDomain()
nz = 137
a = np.random.default_rng(0).normal(250, 30, (nz, ci.getInY(), ci.getInX())).astype(np.float32)
n = nz * 4_000_000
def run(out, label):
    ci.interpolateValues(a, out=out)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); ci.interpolateValues(a, out=out); ts.append(time.perf_counter() - t0)
    print(f"{label}: {min(ts)*1e3:8.2f} ms per 137-level call = {n*4/min(ts)/1e9:5.1f} GB/s, {n/min(ts):.3e} values/s", flush=True)
run(np.empty(n, dtype=np.float32), "pageable output (numpy)          ")
lib = fb.load()
t0 = time.perf_counter(); p = lib.fb200_host_alloc(n * 4); t1 = time.perf_counter()
print(f"fb200_host_alloc of {n*4/1e9:.2f} GB (first time, pages get locked): {(t1-t0)*1e3:.1f} ms")
out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,))
run(out, "fb200_host_alloc output (pinned)  ")
lib.fb200_host_free(p)
t0 = time.perf_counter(); p2 = lib.fb200_host_alloc(n * 4); t1 = time.perf_counter()
print(f"fb200_host_alloc again (recycled, same block: {p2 == p}): {(t1-t0)*1e3:.3f} ms")
lib.fb200_host_free(p2)
