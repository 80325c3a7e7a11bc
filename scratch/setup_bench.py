"""throughput of the once-per-grid (fp64) setup kernels on the BASELINE geometries: points per second"""
import time, numpy as np, torch
import fimex_b200 as fb
from fimex_b200 import Method
SRC = "+proj=latlong +a=6371000 +e=0 +no_defs"
ROT = "+proj=ob_tran +o_proj=longlat +lon_0=-40 +o_lat_p=22 +R=6.371e+06 +no_defs"
STERE = "+proj=stere +lat_0=90 +lon_0=0 +lat_ts=60 +a=6371000 +e=0"
LCC = "+proj=lcc +lat_0=63 +lon_0=15 +lat_1=63 +lat_2=63 +no_defs +R=6.371e+06"
lon = np.arange(1440) * 0.25; lat = 90.0 - np.arange(721) * 0.25
ax2 = (np.arange(2000) - 999.5) * 0.0225
ax4 = -3748750.0 + 2500.0 * np.arange(3000)
def best(fn, reps=3):
    fn(); torch.cuda.synchronize(); b = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); b = min(b, time.perf_counter() - t0)
    return b
torch.zeros(1, device="cuda")
rows = []
for name, m in (("nearestneighbor", Method.NEAREST_NEIGHBOR), ("bilinear", Method.BILINEAR), ("bicubic", Method.BICUBIC)):
    t = best(lambda: fb.CachedInterpolation.fromProjection(m, ROT, ax2, ax2, True, True, SRC, lon, lat, True).createReducedDomain())
    rows.append((f"config 2 tables ({name}): project_axes + 2x points2position + crop + table compile, 4.0e6 targets", t, 4e6))
t = best(lambda: fb.CachedInterpolation.fromProjection(Method.BICUBIC, STERE, ax4, ax4, False, False, SRC, lon, lat, True).createReducedDomain())
rows.append(("config 4 tables (bicubic, polar stereographic), 9.0e6 targets", t, 9e6))
t = best(lambda: fb.CachedVectorReprojection.fromProjection(fb.MIFI_VECTOR_KEEP_SIZE, SRC, STERE, ax4, ax4, fb.PROJ_AXIS, fb.PROJ_AXIS))
rows.append(("config 4 rotation matrix (mifi_get_vector_reproject_matrix: 3 projections + angles), 9.0e6 points", t, 9e6))
lo2, la2 = np.meshgrid(lon, lat)
t = best(lambda: fb.CachedInterpolation.fromCoordinates(Method.COORD_NN, ROT, ax2, ax2, True, True, lo2.ravel(), la2.ravel(), 1440, 721), reps=2)
rows.append(("config 3 coord_nearestneighbor search (1.04e6 sources), 4.0e6 targets", t, 4e6))
t = best(lambda: fb.CachedInterpolation.fromCoordinates(Method.COORD_NN_KD, ROT, ax2, ax2, True, True, lo2.ravel(), la2.ravel(), 1440, 721, maxDistance=30e3), reps=2)
rows.append(("coord_kdtree search, distance of interest 30 km (1.04e6 sources), 4.0e6 targets", t, 4e6))
x = np.radians(np.random.default_rng(1).uniform(-20, 20, 4_000_000)); y = np.radians(np.random.default_rng(2).uniform(-20, 20, 4_000_000))
t = best(lambda: fb.mifi_project_values(ROT, SRC, x, y))
rows.append(("mifi_project_values rotated pole -> lat/long, host arrays (includes H2D + D2H of 64 MB)", t, 4e6))
for name, t, n in rows:
    print(f"{t*1e3:9.2f} ms  {n/t/1e6:9.1f} Mpoints/s  {name}")
