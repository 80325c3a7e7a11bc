"""Seeded differential fuzz of the gathers against the oracle: random source / target sizes, zoom factors from 1/50 (a target
tile sees thousands of source cells: the many-taps and direct-fallback paths of the staged kernels) to 30 (one source cell
covers a whole tile), rotations, targets hanging over every edge of the source grid, level counts around the batch and chunk
sizes, undefined values, scalar / vector / typed slice calls.  Everything is compared bit for bit."""
import os

import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu

import fimex_b200 as fb  # noqa: E402
from fimex_b200 import Method  # noqa: E402


def _case(seed):
    rng = np.random.default_rng(seed)
    inX, inY = int(rng.integers(2, 260)), int(rng.integers(2, 200))
    outX, outY = int(rng.integers(1, 330)), int(rng.integers(1, 140))
    inZ = int(rng.choice([1, 2, 7, 8, 9, 16, 31, 64, 65, 70]))
    zoom = float(np.exp(rng.uniform(np.log(0.02), np.log(30.0))))
    ang = np.radians(rng.uniform(-180, 180))
    cx, cy = rng.uniform(-0.2, 1.2) * inX, rng.uniform(-0.2, 1.2) * inY
    jj, ii = np.meshgrid(np.arange(outX, dtype=np.float64), np.arange(outY, dtype=np.float64))
    u, v = (jj - outX / 2) / zoom, (ii - outY / 2) / zoom
    px = (cx + np.cos(ang) * u - np.sin(ang) * v).ravel()
    py = (cy + np.sin(ang) * u + np.cos(ang) * v).ravel()
    k = rng.integers(0, px.size, max(1, px.size // 20))
    px[k] = np.round(px[k] * 2) / 2  # grid hits and half-cell ties
    py[k[::2]] = np.round(py[k[::2]] * 2) / 2
    field = rng.normal(250, 30, (inZ, inY, inX)).astype(np.float32)
    field[rng.random(field.shape) < rng.choice([0.0, 0.01, 0.2])] = np.nan
    return rng, inX, inY, inZ, outX, outY, px, py, field


@pytest.mark.parametrize("seed", range(int(os.environ.get("FIMEX_B200_FUZZ_SEEDS", "60"))))  # more seeds: a one-off soak run
def test_fuzz_scalar_and_vector(oracle, seed):
    rng, inX, inY, inZ, outX, outY, px, py, field = _case(1000 + seed)
    method = [Method.NEAREST_NEIGHBOR, Method.BILINEAR, Method.BICUBIC][seed % 3]
    ci = fb.CachedInterpolation("x", "y", method, px, py, inX, inY, outX, outY)
    want = oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, field)
    got = ci.interpolateValues(field)
    assert_bit_equal(got, want, f"seed {seed} {method} {inX}x{inY}x{inZ} -> {outX}x{outY}", nan_payload=(method == Method.NEAREST_NEIGHBOR))
    if seed % 2 == 0:  # both components + rotation
        v = rng.normal(0, 8, field.shape).astype(np.float32)
        phi = rng.uniform(-np.pi, np.pi, outX * outY)
        matrix = np.stack([np.cos(phi), np.sin(phi), -np.sin(phi), phi], axis=1).ravel()
        vi = oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, v)
        wu, wv = oracle.vector_reproject_by_matrix(matrix, want, vi, outX, outY, inZ)
        cvr = fb.CachedVectorReprojection(fb.MIFI_VECTOR_KEEP_SIZE, matrix, outX, outY)
        gu, gv = ci.interpolateVector(field, v, cvr)
        assert_bit_equal(gu, wu.reshape(gu.shape), f"seed {seed} u")
        assert_bit_equal(gv, wv.reshape(gv.shape), f"seed {seed} v")
    else:  # the whole slice body with a typed variable
        dt = [np.int16, np.float32, np.int32, np.uint8, np.float64][(seed // 2) % 5]
        fill = fb.default_fill_value(dt)
        if np.dtype(dt).kind == "f":
            data = field.astype(dt)
            data[np.isnan(field)] = fill
        else:
            info = np.iinfo(dt)
            data = np.clip(np.nan_to_num(field, nan=0.0) - 250, max(info.min, -120), min(info.max, 120)).astype(dt)
            data[np.isnan(field)] = np.dtype(dt).type(fill)
        interp = oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, oracle.as_float(data, fill))
        w = oracle.from_float(interp, fill, dt).reshape(inZ, outY, outX)
        g = ci.getDataSlice(data, fill)
        assert g.dtype == np.dtype(dt)
        assert np.array_equal(g.view(np.uint8), w.view(np.uint8)), f"seed {seed} getDataSlice {np.dtype(dt)}"


def _random_crs(rng):
    """one of the four CRS families of the path with random parameters and a random earth figure"""
    fig = rng.choice(["+a=6371000 +e=0", "+R=6.371e+06", "+ellps=sphere", "+a=6378137 +rf=298.257223563", "+a=6378137 +b=6356752.3142",
                      "+datum=WGS84"])
    kind = rng.integers(0, 4)
    if kind == 0:
        return f"+proj=latlong {fig} +no_defs", "ll"
    if kind == 1:  # rotated pole is spherical by construction
        sph = rng.choice(["+a=6371000 +e=0", "+R=6.371e+06", "+ellps=sphere"])
        return (f"+proj=ob_tran +o_proj=longlat +lon_0={rng.uniform(-180, 180):.4f} +o_lat_p={rng.uniform(5, 85):.4f} "
                f"+o_lon_p={rng.choice([0, 0, 10.5, -170])} {sph} +no_defs"), "rot"
    if kind == 2:
        lat0 = rng.choice([90, 90, -90, 60.0, 0.0, -37.5])
        ts = f"+lat_ts={rng.uniform(40, 80) * np.sign(lat0 if lat0 else 1):.3f}" if abs(lat0) == 90 and rng.random() < 0.7 else \
            f"+k={rng.uniform(0.9, 1.0):.5f}"
        return (f"+proj=stere +lat_0={lat0} +lon_0={rng.uniform(-180, 180):.3f} {ts} +x_0={rng.choice([0, 0, 7000, -2.5e5])} "
                f"+y_0={rng.choice([0, 109000.0])} {fig} +no_defs"), "m"
    lat1 = rng.uniform(20, 70) * rng.choice([1, 1, -1])
    lat2 = lat1 if rng.random() < 0.4 else lat1 + rng.uniform(-15, 15)
    return (f"+proj=lcc +lat_1={lat1:.3f} +lat_2={lat2:.3f} +lat_0={lat1 + rng.uniform(-5, 5):.3f} +lon_0={rng.uniform(-180, 180):.3f} "
            f"{fig} +no_defs"), "m"


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_projections_within_1e9_degree(oracle, seed):
    """GPU fp64 forward / inverse projections + pj_transform pipeline (datum shifts included) against the restated PROJ.4 on
    the CPU: <= 1e-9 degree of arc (metric output: that arc on the sphere), identical failure masks, and round trips"""
    rng = np.random.default_rng(5000 + seed)
    src, sk = _random_crs(rng)
    dst, dk = _random_crs(rng)
    n = 3000
    lon = np.radians(rng.uniform(-179, 179, n))
    lat = np.radians(rng.uniform(-88, 88, n))
    ll = "+proj=latlong +a=6371000 +e=0 +no_defs"
    # start from geographic coordinates, go to the source CRS with the ORACLE, then src -> dst with both
    rc, sx, sy = oracle.project_values(ll, src, lon, lat)
    assert rc == 1
    ok = np.isfinite(sx) & np.isfinite(sy)
    sx, sy = sx[ok], sy[ok]
    rc1, wx, wy = oracle.project_values(src, dst, sx, sy)
    rc2, gx, gy = fb.mifi_project_values(src, dst, sx, sy)
    assert rc1 == rc2 == 1, (src, dst)
    fin = np.isfinite(wx) & np.isfinite(wy)
    assert np.array_equal(fin, np.isfinite(gx) & np.isfinite(gy)), (src, dst)
    tol_rad = 1e-9 * np.pi / 180
    if dk == "m":
        # a metric CRS magnifies an arc by its local scale factor (large far from a stereographic centre): compare in the
        # geographic domain by sending both results back through the oracle
        _, bwx, bwy = oracle.project_values(dst, ll, wx[fin], wy[fin])
        _, bgx, bgy = oracle.project_values(dst, ll, gx[fin], gy[fin])
        good = np.isfinite(bwx) & np.isfinite(bgx)
        dx = np.abs(bwx[good] - bgx[good])
        dx = np.minimum(dx, np.abs(dx - 2 * np.pi)) * np.cos(bwy[good])
        dy = np.abs(bwy[good] - bgy[good])
        assert dx.max(initial=0) <= 4 * tol_rad and dy.max(initial=0) <= 4 * tol_rad, (src, dst, dx.max(initial=0), dy.max(initial=0))
        rel = np.abs(gx[fin] - wx[fin]) + np.abs(gy[fin] - wy[fin])
        assert (rel <= 1e-9 * (1 + np.abs(wx[fin]) + np.abs(wy[fin]))).all(), (src, dst)
    else:
        dx = np.abs(gx[fin] - wx[fin])
        dx = np.minimum(dx, np.abs(dx - 2 * np.pi)) * np.cos(np.clip(wy[fin], -np.pi / 2, np.pi / 2))
        dy = np.abs(gy[fin] - wy[fin])
        assert dx.max(initial=0) <= tol_rad and dy.max(initial=0) <= tol_rad, (src, dst, dx.max(initial=0), dy.max(initial=0))
