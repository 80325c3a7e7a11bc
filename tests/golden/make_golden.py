#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libmifi_ref.so = the reference's
src/interpolation.c compiled unmodified, see oracle/Makefile).  Run in the build container, where
/root/reference exists:

    make -C oracle all ref && python tests/golden/make_golden.py

The .npz files are committed; they travel to the GPU box, where /root/reference does not exist.
Projection arithmetic inside these vectors comes from oracle/pj_oracle.c (PROJ is absent from the image),
so vectors that involve a projection pin the reference's *use* of PROJ, not PROJ itself.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import (BICUBIC, BILINEAR, LATITUDE, LONGITUDE, NEAREST_NEIGHBOR, PROJ_AXIS, Oracle, Reference)  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

EMEP = "+ellps=sphere +a=127.4 +e=0 +proj=stere +lat_0=90 +lon_0=-32 +lat_ts=60 +x_0=7 +y_0=109"
LATLONG = "+ellps=sphere +a=6370 +e=0 +proj=latlong"
SRC_LL = "+proj=latlong +a=6371000 +e=0 +no_defs"
ROTPOLE = "+proj=ob_tran +o_proj=longlat +lon_0=-40 +o_lat_p=22 +R=6.371e+06 +no_defs"
STERE = "+proj=stere +lat_0=90 +lon_0=0 +lat_ts=60 +a=6371000 +e=0"
LCC = "+proj=lcc +lat_0=63 +lon_0=15 +lat_1=63 +lat_2=63 +no_defs +R=6.371e+06"
WGS84 = "+proj=latlong +datum=WGS84 +towgs84=0,0,0 +no_defs"


def adversarial_positions(rng, ix, iy, n):
    """random + edge/half-cell/integer positions for an ix*iy grid"""
    x = rng.uniform(-1.5, ix + 0.5, n)
    y = rng.uniform(-1.5, iy + 0.5, n)
    special_x = np.array([-1.0, -0.5, -0.5 + 1e-12, -0.25, 0.0, 0.5, 1.0, ix - 2.0, ix - 1.5, ix - 1.0, ix - 1.0 + 1e-9, ix - 0.75,
                          ix - 0.5, ix - 0.5 - 1e-12, ix, -999.0, 2.5, 3.0, 1.0 - 1e-15])
    special_y = np.array([-1.0, -0.5, -0.5 + 1e-12, -0.25, 0.0, 0.5, 1.0, iy - 2.0, iy - 1.5, iy - 1.0, iy - 1.0 + 1e-9, iy - 0.75,
                          iy - 0.5, iy - 0.5 - 1e-12, iy, -999.0, 2.5, 3.0, 1.0 - 1e-15])
    gx, gy = np.meshgrid(special_x, special_y)
    x = np.concatenate([x, gx.ravel(), rng.integers(0, ix, 64).astype(float)])
    y = np.concatenate([y, gy.ravel(), rng.integers(0, iy, 64).astype(float)])
    return x, y


def main():
    ref = Reference()
    orc = Oracle()
    rng = np.random.default_rng(20261018)

    # ---- 1. EMEP country map (reference test/testInterpolation.cc:280-393; data test/inData.txt) -----
    d = np.loadtxt("/root/reference/test/inData.txt")
    emep = np.full((150, 170), np.nan, dtype=np.float32)
    emep[d[:, 1].astype(int) - 1, d[:, 0].astype(int) - 1] = d[:, 2]
    lon = (np.arange(180) + 1) / 2.0 - 30
    lat = (np.arange(90) + 1) / 2.0 + 30
    outs = {}
    for name, m in (("nn", NEAREST_NEIGHBOR), ("bilinear", BILINEAR), ("bicubic", BICUBIC)):
        rc, o = ref.interpolate_f(m, EMEP, emep, np.arange(170) + 1.0, np.arange(150) + 1.0, PROJ_AXIS, PROJ_AXIS, 1, LATLONG, lon,
                                  lat, LONGITUDE, LATITUDE)
        assert rc == 1 and o[0, 25, 9] == 32.0
        outs[name] = o
    np.savez_compressed(os.path.join(OUT, "emep.npz"), infield=emep, lon=lon, lat=lat, **outs)

    # ---- 2. per-point kernels on random + adversarial positions ---------------------------------------
    ix, iy, iz = 23, 17, 3
    field = rng.normal(250, 30, (iz, iy, ix)).astype(np.float32)
    field[rng.random(field.shape) < 0.03] = np.nan
    field[0, 0, 0] = np.inf
    field[1, 5, 5] = -0.0
    px, py = adversarial_positions(rng, ix, iy, 3000)
    ub = np.array([orc.bilinear_is_ub(a, b, ix, iy) for a, b in zip(px, py)])
    res = {}
    for name, m in (("nn", NEAREST_NEIGHBOR), ("bilinear", BILINEAR), ("bicubic", BICUBIC)):
        o = np.empty((iz, px.size), dtype=np.float32)
        for i in range(px.size):
            if m == BILINEAR and ub[i]:
                o[:, i] = np.nan  # the reference reads out of bounds here (interpolation.c:936); masked
            else:
                o[:, i] = ref.get_values(m, field, px[i], py[i], ix, iy, iz)
        res[name] = o
    np.savez_compressed(os.path.join(OUT, "kernels.npz"), field=field, px=px, py=py, ub=ub, **res)

    # ---- 3. points2position ------------------------------------------------------------------------------
    axes = {
        "asc": (np.array([1.0, 2, 3, 4, 5]), PROJ_AXIS),
        "desc": (np.array([5.0, 4, 3, 2, 1]), PROJ_AXIS),
        "lon360": (np.radians(np.arange(1440) * 0.25), LONGITUDE),
        "lon180": (np.radians(-180 + np.arange(720) * 0.5), LONGITUDE),
        "lon_regional": (np.radians(5.0 + np.arange(17) * 0.1), LONGITUDE),
        "lon_desc": (np.radians(359.0 - np.arange(360)), LONGITUDE),
        "lat_desc": (np.radians(90 - np.arange(721) * 0.25), LATITUDE),
        "nonuniform": (np.cumsum(rng.uniform(0.5, 2.0, 40)), PROJ_AXIS),
        "metric": (-3748750.0 + 2500.0 * np.arange(3000), PROJ_AXIS),
    }
    p2p = {}
    for k, (ax, t) in axes.items():
        lo, hi = ax.min(), ax.max()
        span = hi - lo
        pts = rng.uniform(lo - 0.3 * span, hi + 0.3 * span, 500)
        if t == LONGITUDE:
            pts = np.concatenate([pts, rng.uniform(-2 * np.pi, 2 * np.pi, 300), [np.pi, -np.pi, 0.0, 2 * np.pi, -1e-3]])
        pts = np.concatenate([pts, ax[::7], [np.nan, np.inf, -np.inf, 1e300]])
        p2p[k + "_axis"] = ax
        p2p[k + "_type"] = np.array(t)
        p2p[k + "_in"] = pts
        p2p[k + "_out"] = ref.points2position(pts, ax, t)
    np.savez_compressed(os.path.join(OUT, "points2position.npz"), **p2p)

    # ---- 4. projections + vector matrices (PROJ arithmetic = oracle/pj_oracle.c) ----------------------
    proj = {}
    deg_axes_x = np.radians(np.linspace(-22.4875, 22.4875, 9))
    deg_axes_y = np.radians(np.linspace(-22.4875, 22.4875, 7))
    m_axes_x = np.linspace(-3748750.0, 3748750.0, 9)
    m_axes_y = np.linspace(-3748750.0, 3748750.0, 7)
    lcc_x = np.linspace(-1.5e6, 1.5e6, 9)
    lcc_y = np.linspace(-3.0e6, 3.0e6, 7)
    cases = {
        "rot_to_ll": (ROTPOLE, SRC_LL, deg_axes_x, deg_axes_y),
        "stere_to_ll": (STERE, SRC_LL, m_axes_x, m_axes_y),
        "lcc_to_ll": (LCC, SRC_LL, lcc_x, lcc_y),
        "lcc_to_wgs84": (LCC, WGS84, lcc_x, lcc_y),
        "emep_to_ll": (EMEP, LATLONG, np.array([6.0, 7, 8]), np.array([108.0, 109, 110])),
        "ll_to_rot": (SRC_LL, ROTPOLE, np.radians(np.linspace(-60, 30, 9)), np.radians(np.linspace(20, 80, 7))),
        "ll_to_stere": (SRC_LL, STERE, np.radians(np.linspace(-180, 180, 9)), np.radians(np.linspace(30, 90, 7))),
        "ll_to_lcc": (SRC_LL, LCC, np.radians(np.linspace(-20, 50, 9)), np.radians(np.linspace(40, 85, 7))),
        "rot_to_stere": (ROTPOLE, STERE, deg_axes_x, deg_axes_y),
    }
    for k, (pin, pout, xa, ya) in cases.items():
        rc, xo, yo = ref.project_axes(pin, pout, xa, ya)
        assert rc == 1, k
        proj[k + "_xa"], proj[k + "_ya"], proj[k + "_xo"], proj[k + "_yo"] = xa, ya, xo, yo
    np.savez_compressed(os.path.join(OUT, "projections.npz"), **proj)

    vec = {}
    vcases = {
        "ll_rot": (SRC_LL, ROTPOLE, np.linspace(-22.4875, 22.4875, 9), np.linspace(-22.4875, 22.4875, 7), LONGITUDE, LATITUDE),
        "ll_stere": (SRC_LL, STERE, m_axes_x, m_axes_y, PROJ_AXIS, PROJ_AXIS),
        "ll_lcc": (SRC_LL, LCC, lcc_x, lcc_y, PROJ_AXIS, PROJ_AXIS),
        "stere_ll": (STERE, SRC_LL, np.linspace(-30, 40, 9), np.linspace(50, 85, 7), LONGITUDE, LATITUDE),
        "rot_stere": (ROTPOLE, STERE, m_axes_x / 4, m_axes_y / 4, PROJ_AXIS, PROJ_AXIS),
    }
    for k, (pin, pout, xa, ya, xt, yt) in vcases.items():
        rc, m = ref.vector_matrix(pin, pout, xa, ya, xt, yt)
        assert rc == 1, k
        u = rng.normal(0, 10, (2, ya.size, xa.size)).astype(np.float32)
        v = rng.normal(0, 10, (2, ya.size, xa.size)).astype(np.float32)
        u[0, 1, 1] = np.nan
        ur, vr = ref.vector_reproject_by_matrix(m, u, v, xa.size, ya.size, 2)
        vec[k + "_xa"], vec[k + "_ya"], vec[k + "_m"] = xa, ya, m
        vec[k + "_u"], vec[k + "_v"], vec[k + "_ur"], vec[k + "_vr"] = u, v, ur.reshape(u.shape), vr.reshape(v.shape)
    np.savez_compressed(os.path.join(OUT, "vectors.npz"), **vec)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
