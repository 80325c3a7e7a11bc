import time, os, numpy as np, torch
import fimex_b200 as fb
from fimex_b200 import Method
SRC = "+proj=latlong +a=6371000 +e=0 +no_defs"
DST = "+proj=ob_tran +o_proj=longlat +lon_0=-40 +o_lat_p=22 +R=6.371e+06 +no_defs"
lon = np.arange(1440) * 0.25; lat = 90.0 - np.arange(721) * 0.25
ax = (np.arange(2000) - 999.5) * 0.0225
ci = fb.CachedInterpolation.fromProjection(Method.BILINEAR, DST, ax, ax, True, True, SRC, lon, lat, True)
ci.createReducedDomain()
h_in = torch.randn((137, ci.getInY(), ci.getInX())).pin_memory()
h_out = torch.empty(137 * 4_000_000, dtype=torch.float32).pin_memory()
a, b = h_in.numpy(), h_out.numpy()
ci.interpolateValues(a, out=b)
ts = []
for _ in range(8):
    t0 = time.perf_counter(); ci.interpolateValues(a, out=b); ts.append(time.perf_counter() - t0)
print(os.environ.get("FIMEX_B200_HOST_CHUNK_MB", "192"), "MB chunks: per call ms", [round(t * 1e3, 2) for t in ts], " -> %.1f GB/s" % (b.nbytes / min(ts) / 1e9), flush=True)
