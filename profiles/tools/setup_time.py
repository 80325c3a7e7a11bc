import time, sys, os
import numpy as np
import fimex_b200 as fb
from fimex_b200 import Method
STERE = "+proj=stere +lat_0=90 +lon_0=0 +lat_ts=60 +a=6371000 +e=0"
SRC_LL = "+proj=latlong +a=6371000 +e=0 +no_defs"
LON = np.arange(1440) * 0.25
LAT = 90.0 - np.arange(721) * 0.25
ax = -3748750.0 + 2500.0 * np.arange(3000)
import torch
torch.zeros(1, device="cuda")
for rep in range(3):
    t0 = time.perf_counter()
    ci = fb.CachedInterpolation.fromProjection(Method.BICUBIC, STERE, ax, ax, False, False, SRC_LL, LON, LAT, True)
    t1 = time.perf_counter()
    ci.createReducedDomain()
    t2 = time.perf_counter()
    print("rep", rep, "fromProjection %.3f s  createReducedDomain %.3f s" % (t1 - t0, t2 - t1), flush=True)
    del ci
