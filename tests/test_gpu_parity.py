"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle and the golden
vectors produced by the compiled reference.  Run on the B200 box with `pytest -m gpu`.

Bars (BASELINE.json north_star): bit-exact for nearest-neighbour indices, copied values and fill masks; bit-exact
as well for bilinear / bicubic / rotation here, because the kernels replay the reference's operation order
without FMA contraction (the stated bar is 1e-5 relative); <= 1e-9 degree for fp64 coordinate transforms.
"""
import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu

import fimex_b200 as fb  # noqa: E402
from fimex_b200 import Method  # noqa: E402

EMEP = "+ellps=sphere +a=127.4 +e=0 +proj=stere +lat_0=90 +lon_0=-32 +lat_ts=60 +x_0=7 +y_0=109"
LATLONG = "+ellps=sphere +a=6370 +e=0 +proj=latlong"
SRC_LL = "+proj=latlong +a=6371000 +e=0 +no_defs"
ROTPOLE = "+proj=ob_tran +o_proj=longlat +lon_0=-40 +o_lat_p=22 +R=6.371e+06 +no_defs"
STERE = "+proj=stere +lat_0=90 +lon_0=0 +lat_ts=60 +a=6371000 +e=0"
LCC = "+proj=lcc +lat_0=63 +lon_0=15 +lat_1=63 +lat_2=63 +no_defs +R=6.371e+06"
WGS84 = "+proj=latlong +datum=WGS84 +towgs84=0,0,0 +no_defs"
DEG = np.pi / 180
TOL_RAD = 1e-9 * DEG  # 1e-9 degree


# =====================================================================================================
# the reference's known-answer tests, through the drop-in mifi_* symbols (test/testInterpolation.cc)
# =====================================================================================================
def test_mifi_points2position():
    # :49-58 and :61-70
    rc, p = fb.mifi_points2position([-3.0, 5.0, 1.3, 2.0, 6.0], [1.0, 2, 3, 4, 5], fb.PROJ_AXIS)
    assert rc == fb.MIFI_OK and np.allclose(p, [-4.0, 4.0, 0.3, 1.0, 5.0], atol=1e-10, rtol=0)
    rc, p = fb.mifi_points2position([-3.0, 5.0, 1.3, 2.0, 6.0], [5.0, 4, 3, 2, 1], fb.PROJ_AXIS)
    assert rc == fb.MIFI_OK and np.allclose(p, [8.0, 0.0, 3.7, 3.0, -1.0], atol=1e-10, rtol=0)


def test_mifi_get_values_f():
    # :73-80
    rc, out = fb.mifi_get_values_f([1.0, 2.0, 1.0, 2.0], 0.3, 0.3, 2, 2, 1)
    assert rc == fb.MIFI_OK and out[0] == 1.0


def test_mifi_get_values_bilinear_f():
    # :83-112
    f = np.array([1.0, 2.0, 2.0, 1 + np.sqrt(np.float32(2.0))], dtype=np.float32)
    g = lambda x, y: float(fb.mifi_get_values_bilinear_f(f, x, y, 2, 2, 1)[1][0])
    assert abs(g(0.3, 0.0) - 1.3) < 1e-6 and abs(g(0.3, 0.0001) - 1.3) < 1e-4
    assert abs(g(0.0, 0.3) - 1.3) < 1e-6 and abs(g(0.0001, 0.3) - 1.3) < 1e-4
    assert not np.isnan(g(0, 0)) and not np.isnan(g(1, 1))
    for x, y in ((1.5, 0.5), (0.5, 1.5), (0.5, -0.5), (-0.5, 0.5)):
        assert np.isnan(g(x, y))


def test_mifi_get_values_bicubic_f():
    # :115-155
    f = np.array([1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1], dtype=np.float32)
    ft = f.reshape(4, 4).T.copy().ravel()
    g = lambda a, x, y: float(fb.mifi_get_values_bicubic_f(a, x, y, 4, 4, 1)[1][0])
    for a, x, y, want in ((f, 1, 1, 2.0), (f, 1, 1.99999, 2.0), (f, 1, 1.5, 2.125), (f, 1.5, 1, 2.0), (ft, 1, 1, 2.0), (ft, 1.99999, 1, 2.0),
                          (ft, 1.5, 1, 2.125), (ft, 1, 1.5, 2.0)):
        assert g(a, x, y) == pytest.approx(want, rel=1e-5)
    for x, y in ((0.5, 1), (1, 0.5), (2.5, 1), (1, 2.5)):
        assert np.isnan(g(f, x, y))


def test_mifi_project_axes_emep():
    # :265-278
    rc, x, y = fb.mifi_project_axes(EMEP, LATLONG, [6.0, 7, 8], [108.0, 109, 110])
    assert rc == fb.MIFI_OK and np.all(y / DEG > 89)
    assert y[4] / DEG == pytest.approx(90.0, abs=1e-9)


def test_mifi_interpolate_f_emep(golden, oracle):
    # :280-393.  The golden fields come from the COMPILED reference, whose projection arithmetic runs with glibc on the host;
    # here it runs with the CUDA math library, so the fractional positions may differ in the last ulps.  The comparison is
    # therefore split so that nothing is waved through:
    #   (1) positions: GPU vs the CPU pipeline (mifi_project_axes + mifi_points2position) within 1e-9 degree of arc;
    #   (2) values, given the GPU's own positions: bit-identical to the oracle's kernels (no tolerance);
    #   (3) golden field: differences only where (1) allows them -- nearest neighbour / NaN masks may flip only at positions
    #       within max|dpos| of a (half-)cell boundary, interpolated values move by at most |dpos| x the local gradient.
    g = golden("emep")
    ax, ay = np.arange(170) + 1.0, np.arange(150) + 1.0
    lon, lat = g["lon"], g["lat"]
    rc, x, y = fb.mifi_project_axes(LATLONG, EMEP, np.radians(lon), np.radians(lat))  # convertAxis: degrees -> radians (:221-229)
    assert rc == fb.MIFI_OK
    rc, gx = fb.mifi_points2position(x, ax, fb.PROJ_AXIS)
    rc, gy = fb.mifi_points2position(y, ay, fb.PROJ_AXIS)
    rc, ox, oy = oracle.project_axes(LATLONG, EMEP, np.radians(lon), np.radians(lat))
    ox, oy = oracle.points2position(ox, ax, 0), oracle.points2position(oy, ay, 0)
    dpos = max(np.abs(gx - ox).max(), np.abs(gy - oy).max())
    assert dpos <= 1e-9 * DEG * 127.4 * 2 + 1e-12, dpos  # 1e-9 degree of arc in grid units (a = 127.4 cells, scale <= 2 at the pole)
    eps = dpos + 1e-12
    near_half = (np.abs(gx - np.floor(gx) - 0.5) <= eps) | (np.abs(gy - np.floor(gy) - 0.5) <= eps)
    near_int = (np.abs(gx - np.round(gx)) <= eps) | (np.abs(gy - np.round(gy)) <= eps)
    field = g["infield"]
    grad = max(np.nanmax(np.abs(np.diff(field, axis=0))), np.nanmax(np.abs(np.diff(field, axis=1))))
    for name, m in (("nn", Method.NEAREST_NEIGHBOR), ("bilinear", Method.BILINEAR), ("bicubic", Method.BICUBIC)):
        rc, out = fb.mifi_interpolate_f(m, EMEP, field, ax, ay, fb.PROJ_AXIS, fb.PROJ_AXIS, 1,
                                        LATLONG, lon, lat, fb.LONGITUDE, fb.LATITUDE)
        assert rc == fb.MIFI_OK
        assert abs(out[0, 25, 9] - 32) < 1e-6  # the reference's own assertion (:336, :362, :388)
        same_pos = oracle.cached_interpolate(int(m), gx, gy, 170, 150, 180, 90, field[None])
        assert_bit_equal(out, same_pos, f"emep {name} on the GPU's positions", nan_payload=(name == "nn"))
        want = g[name]
        may_flip = (near_half | near_int).reshape(out.shape)
        mask_flips = np.isnan(out) != np.isnan(want)
        assert not (mask_flips & ~may_flip).any(), (name, int(mask_flips.sum()), int(may_flip.sum()))
        both = ~np.isnan(out) & ~np.isnan(want)
        if name == "nn":
            assert not ((out != want) & both & ~may_flip).any(), name
        else:  # 16 taps with weights up to 1.27 for bicubic: |d value| <= 3 * 16 * gradient * |dpos|, plus one ulp of the value
            bound = 48 * grad * eps + np.spacing(np.abs(want[both]).astype(np.float32)) * 2
            assert (np.abs(out[both] - want[both])[~may_flip[both]] <= bound[~may_flip[both]]).all(), name


@pytest.mark.parametrize("lon0,tol", [(90, 1e-4), (180, 1e-5)])
def test_mifi_vector_reproject_values_rotate(lon0, tol):
    # :396-453 and :455-512
    a = "+ellps=sphere +a=127.4 +e=0 +proj=stere +lat_0=90 +lon_0=0 +lat_ts=60"
    b = f"+ellps=sphere +a=127.4 +e=0 +proj=stere +lat_0=90 +lon_0={lon0} +lat_ts=60"
    ax = np.arange(5) - 2.0
    u = np.arange(25, dtype=np.float32)
    v = 25 - np.arange(25, dtype=np.float32)
    _, uo = fb.mifi_interpolate_f(0, a, u, ax, ax, 0, 0, 1, b, ax, ax, 0, 0)
    _, vo = fb.mifi_interpolate_f(0, a, v, ax, ax, 0, 0, 1, b, ax, ax, 0, 0)
    rc, ur, vr = fb.mifi_vector_reproject_values_f(fb.MIFI_VECTOR_KEEP_SIZE, a, b, uo, vo, ax, ax, 0, 0, 1)
    assert rc == fb.MIFI_OK
    uo, vo, ur, vr = uo.ravel(), vo.ravel(), ur.ravel(), vr.ravel()
    if lon0 == 90:
        assert np.all(np.abs(vo - ur) < tol) and np.all(np.abs(uo + vr) < tol)
    else:
        assert np.all(np.abs(vo + vr) < tol) and np.all(np.abs(uo + ur) < tol)


def test_mifi_vector_reproject_keep_size_and_directions():
    # :515-583
    ai, aj = np.arange(4) + 6.0, np.arange(4) + 108.0
    lon, lat = np.arange(4) * 60.0, np.arange(4) / 2.0 + 88.5
    u = np.arange(16, dtype=np.float32)
    v = -16 + np.arange(16, dtype=np.float32)
    _, uo = fb.mifi_interpolate_f(0, EMEP, u, ai, aj, 0, 0, 1, LATLONG, lon, lat, fb.LONGITUDE, fb.LATITUDE)
    _, vo = fb.mifi_interpolate_f(0, EMEP, v, ai, aj, 0, 0, 1, LATLONG, lon, lat, fb.LONGITUDE, fb.LATITUDE)
    rc, ur, vr = fb.mifi_vector_reproject_values_f(0, EMEP, LATLONG, uo, vo, lon, lat, fb.LONGITUDE, fb.LATITUDE, 1)
    assert rc == fb.MIFI_OK
    d = ur.ravel().astype(np.float64)**2 + vr.ravel().astype(np.float64)**2 - uo.ravel().astype(np.float64)**2 - vo.ravel().astype(np.float64)**2
    d = d[~np.isnan(d)]
    assert d.size > 0 and np.all(np.abs(d) < 1e-3)
    # :586-654
    a = "+ellps=sphere +a=127.4 +e=0 +proj=stere +lat_0=90 +lon_0=0 +lat_ts=60"
    ax = (np.arange(5) - 2) * 1000.0
    xf, yf = np.meshgrid(ax, ax)
    rc, m = fb.mifi_get_vector_reproject_matrix_field(a, LATLONG, xf.ravel(), yf.ravel(), 5, 5)
    assert rc == fb.MIFI_OK
    rc, ang = fb.mifi_vector_reproject_direction_by_matrix_f(0, m, np.zeros(25, dtype=np.float32), 5, 5, 1)
    close = lambda want, got: abs(got - want) <= 0.01 * max(abs(want), abs(got))
    assert close(315, ang[0]) and close(270, ang[10]) and close(225, ang[20])
    assert close(180, ang[2 + 15]) and close(180, ang[2 + 20])
    assert close(45, ang[4]) and close(90, ang[4 + 10]) and close(135, ang[4 + 20])


# =====================================================================================================
# golden vectors from the compiled reference: identical inputs => identical bits
# =====================================================================================================
def test_golden_kernels_bit_exact(golden):
    g = golden("kernels")
    field, px, py = g["field"], g["px"], g["py"]
    iz, iy, ix = field.shape
    n = px.size
    for name, m in (("nn", Method.NEAREST_NEIGHBOR), ("bilinear", Method.BILINEAR), ("bicubic", Method.BICUBIC)):
        ci = fb.CachedInterpolation("x", "y", m, px, py, ix, iy, n, 1)
        out = ci.interpolateValues(field).reshape(iz, n)
        assert_bit_equal(out, g[name], f"{name} vs compiled reference", nan_payload=(name == "nn"))


def test_golden_points2position_bit_exact(golden):
    g = golden("points2position")
    for k in ("asc", "desc", "lon360", "lon180", "lon_regional", "lon_desc", "lat_desc", "nonuniform", "metric"):
        rc, got = fb.mifi_points2position(g[k + "_in"], g[k + "_axis"], int(g[k + "_type"]))
        assert rc == fb.MIFI_OK
        assert np.array_equal(got.view(np.uint64), g[k + "_out"].view(np.uint64)), k


def test_golden_projections_within_1e9_degree(golden):
    g = golden("projections")
    cases = {
        "rot_to_ll": (ROTPOLE, SRC_LL, True), "stere_to_ll": (STERE, SRC_LL, True), "lcc_to_ll": (LCC, SRC_LL, True),
        "lcc_to_wgs84": (LCC, WGS84, True), "emep_to_ll": (EMEP, LATLONG, True), "ll_to_rot": (SRC_LL, ROTPOLE, True),
        "ll_to_stere": (SRC_LL, STERE, False), "ll_to_lcc": (SRC_LL, LCC, False), "rot_to_stere": (ROTPOLE, STERE, False),
    }
    for k, (pin, pout, angular) in cases.items():
        rc, xo, yo = fb.mifi_project_axes(pin, pout, g[k + "_xa"], g[k + "_ya"])
        assert rc == fb.MIFI_OK, k
        wx, wy = g[k + "_xo"], g[k + "_yo"]
        finite = np.isfinite(wx) & np.isfinite(wy)
        assert np.array_equal(finite, np.isfinite(xo) & np.isfinite(yo)), k
        tol = TOL_RAD if angular else TOL_RAD * 6371000.0  # 1e-9 degree of arc
        dx = np.abs(xo[finite] - wx[finite])
        if angular:
            dx = np.minimum(dx, np.abs(dx - 2 * np.pi))
        assert dx.max() <= tol and np.abs(yo[finite] - wy[finite]).max() <= tol, (k, dx.max())


def test_golden_vector_rotation(golden):
    g = golden("vectors")
    cases = {"ll_rot": (SRC_LL, ROTPOLE, fb.LONGITUDE, fb.LATITUDE), "ll_stere": (SRC_LL, STERE, 0, 0), "ll_lcc": (SRC_LL, LCC, 0, 0),
             "stere_ll": (STERE, SRC_LL, fb.LONGITUDE, fb.LATITUDE), "rot_stere": (ROTPOLE, STERE, 0, 0)}
    for k, (pin, pout, xt, yt) in cases.items():
        m, u, v = g[k + "_m"], g[k + "_u"], g[k + "_v"]
        oz, oy, ox = u.shape
        # rotation with the reference's matrix: bit-exact
        rc, ur, vr = fb.mifi_vector_reproject_values_by_matrix_f(0, m, u, v, ox, oy, oz)
        assert rc == fb.MIFI_OK
        assert_bit_equal(ur.reshape(u.shape), g[k + "_ur"], k + " u")
        assert_bit_equal(vr.reshape(v.shape), g[k + "_vr"], k + " v")
        cvr = fb.CachedVectorReprojection(fb.MIFI_VECTOR_KEEP_SIZE, m, ox, oy)
        u2, v2 = u.copy(), v.copy()
        cvr.reprojectValues(u2, v2)
        assert_bit_equal(u2, g[k + "_ur"], k + " u (class)")
        # matrix built on the device: angle within 1e-9 degree of the reference's
        rc, m2 = fb.mifi_get_vector_reproject_matrix(pin, pout, g[k + "_xa"], g[k + "_ya"], xt, yt)
        assert rc == fb.MIFI_OK
        dphi = np.abs(m2[3::4] - m[3::4])
        assert dphi.max() <= 1e-7 * DEG, (k, dphi.max())  # angle from differences of 1e-3-cell steps: conditioned ~1e2
        assert np.abs(m2[0::4] - m[0::4]).max() < 1e-9 and np.abs(m2[1::4] - m[1::4]).max() < 1e-9
        assert np.array_equal(m2[2::4], -m2[1::4])


# =====================================================================================================
# seeded random parity against the oracle: CachedInterpolation (A3-A8)
# =====================================================================================================
def _random_case(seed, inX, inY, inZ, outX, outY, nan_frac=0.02, spill=1.5):
    rng = np.random.default_rng(seed)
    field = rng.normal(250, 30, (inZ, inY, inX)).astype(np.float32)
    field[rng.random(field.shape) < nan_frac] = np.nan
    n = outX * outY
    px = rng.uniform(-spill, inX - 1 + spill, n)
    py = rng.uniform(-spill, inY - 1 + spill, n)
    k = n // 10
    px[:k] = np.round(px[:k] * 2) / 2
    py[k // 2:k + k // 2] = np.round(py[k // 2:k + k // 2] * 2) / 2
    px[-3:] = [-999.0, 0.0, inX - 1.0]
    py[-3:] = [-999.0, inY - 1.0, 0.0]
    return field, px, py


def _mask_ub(oracle, method, px, py, inX, inY, out):
    """positions where the reference itself reads out of bounds (interpolation.c:936) are NaN on both sides"""
    if method != Method.BILINEAR:
        return 0
    ub = np.array([oracle.bilinear_is_ub(a, b, inX, inY) for a, b in zip(px, py)])
    return int(ub.sum())


@pytest.mark.parametrize("method", [Method.NEAREST_NEIGHBOR, Method.BILINEAR, Method.BICUBIC])
@pytest.mark.parametrize("shape", [(37, 23, 3, 40, 25), (64, 48, 5, 101, 7), (9, 7, 70, 33, 4), (200, 150, 2, 256, 128), (2, 2, 1, 5, 5)])
def test_cached_interpolation_bit_exact(oracle, method, shape):
    inX, inY, inZ, outX, outY = shape
    field, px, py = _random_case(hash(shape) % 10000 + int(method), inX, inY, inZ, outX, outY)
    want = oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, field)
    ci = fb.CachedInterpolation("x", "y", method, px, py, inX, inY, outX, outY)
    assert (ci.getInX(), ci.getInY(), ci.getOutX(), ci.getOutY()) == (inX, inY, outX, outY)
    got = ci.interpolateValues(field)
    assert got.shape == (inZ, outY, outX)
    assert_bit_equal(got, want, f"method {method} shape {shape}", nan_payload=(method == Method.NEAREST_NEIGHBOR))
    _mask_ub(oracle, method, px, py, inX, inY, got)


def _smooth_positions(inX, inY, outX, outY, angle_deg, zoom, seed):
    """target grid = rotated, finer copy of the source grid (several target points per source cell), slightly warped"""
    rng = np.random.default_rng(seed)
    a = np.radians(angle_deg)
    jj, ii = np.meshgrid(np.arange(outX, dtype=np.float64), np.arange(outY, dtype=np.float64))
    u, v = (jj - outX / 2) / zoom, (ii - outY / 2) / zoom
    px = inX / 2 + np.cos(a) * u - np.sin(a) * v + 0.3 * np.sin(v / 7)
    py = inY / 2 + np.sin(a) * u + np.cos(a) * v + 0.3 * np.cos(u / 5)
    px, py = px.ravel(), py.ravel()
    k = rng.integers(0, px.size, 50)
    px[k] = np.round(px[k])  # exact grid hits: zero weights, NaN taps still poison (trap 1)
    return px, py


@pytest.mark.parametrize("method", [Method.BICUBIC, Method.BILINEAR, Method.NEAREST_NEIGHBOR])
@pytest.mark.parametrize("inX,inY,inZ,outX,outY,angle,zoom", [
    (60, 50, 19, 200, 150, 17.0, 5.0),    # fast path: few taps per tile, partial batches (19 = 2*8 + 3), whole tiles
    (60, 50, 70, 203, 77, -33.0, 6.5),    # outX not a multiple of 4 (scalar stores), partial tiles, two level chunks
    (300, 40, 9, 256, 96, 3.0, 0.6),      # target coarser than source: hundreds of taps per tile (one-level batches)
    (1500, 8, 3, 64, 48, 80.0, 0.02),     # thousands of taps per tile: direct fallback inside the staged kernel
])
def test_staged_gathers_on_structured_grids(oracle, method, inX, inY, inZ, outX, outY, angle, zoom):
    """The staged (shared-memory) kernels take their fast paths only on structured grids; random positions
    (test_cached_interpolation_bit_exact) exercise their many-taps paths."""
    px, py = _smooth_positions(inX, inY, outX, outY, angle, zoom, 3)
    rng = np.random.default_rng(8)
    field = rng.normal(250, 30, (inZ, inY, inX)).astype(np.float32)
    field[rng.random(field.shape) < 0.01] = np.nan
    ci = fb.CachedInterpolation("x", "y", method, px, py, inX, inY, outX, outY)
    got = ci.interpolateValues(field)
    want = oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, field)
    assert_bit_equal(got, want, f"structured grid, method {method}", nan_payload=(method == Method.NEAREST_NEIGHBOR))
    # both components of a vector in one pass, with and without the rotation
    v = rng.normal(0, 8, field.shape).astype(np.float32)
    phi = rng.uniform(-np.pi, np.pi, outX * outY)
    matrix = np.stack([np.cos(phi), np.sin(phi), -np.sin(phi), phi], axis=1).ravel()
    vi = oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, v)
    wu, wv = oracle.vector_reproject_by_matrix(matrix, want, vi, outX, outY, inZ)
    cvr = fb.CachedVectorReprojection(fb.MIFI_VECTOR_KEEP_SIZE, matrix, outX, outY)
    gu, gv = ci.interpolateVector(field, v, cvr)
    assert_bit_equal(gu, wu.reshape(gu.shape), "fused u")
    assert_bit_equal(gv, wv.reshape(gv.shape), "fused v")
    pu, pv = ci.interpolateVector(field, v, None)
    assert_bit_equal(pu, want, "no-rotation u")
    assert_bit_equal(pv, vi, "no-rotation v")


def test_cached_interpolation_device_path_equals_host_path(oracle):
    import torch
    inX, inY, inZ, outX, outY = 120, 90, 40, 300, 200  # 60000 points: the 128-bit store path; 40 levels
    field, px, py = _random_case(5, inX, inY, inZ, outX, outY)
    for method in (Method.NEAREST_NEIGHBOR, Method.BILINEAR, Method.BICUBIC, Method.COORD_NN):
        ci = fb.CachedInterpolation("x", "y", method, px, py, inX, inY, outX, outY)
        host = ci.interpolateValues(field)
        dev = ci.interpolateValues(torch.from_numpy(field).cuda())
        torch.cuda.synchronize()
        assert_bit_equal(dev.cpu().numpy(), host, f"device vs host path, method {method}")
        want = oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, field)
        assert_bit_equal(host, want, f"oracle, method {method}")


def test_reduced_domain_matches_oracle(oracle):
    inX, inY, inZ, outX, outY = 300, 200, 3, 64, 48
    rng = np.random.default_rng(11)
    px = rng.uniform(100.3, 140.7, outX * outY)
    py = rng.uniform(50.2, 77.9, outX * outY)
    field = rng.normal(0, 1, (inZ, inY, inX)).astype(np.float32)
    red, opx, opy, oinX, oinY, ominX, ominY = oracle.reduced_domain(px, py, inX, inY)
    assert red
    for method in (Method.BILINEAR, Method.BICUBIC, Method.NEAREST_NEIGHBOR):
        ci = fb.CachedInterpolation("x", "y", method, px, py, inX, inY, outX, outY)
        assert ci.createReducedDomain()
        assert (ci.getInX(), ci.getInY()) == (oinX, oinY)
        assert ci.reducedDomain()[2:4] == (ominX, ominY)
        gx, gy = ci.points()
        assert np.array_equal(gx.view(np.uint64), opx.view(np.uint64)) and np.array_equal(gy.view(np.uint64), opy.view(np.uint64))
        cropped = ci.getInputDataSlice(field)
        assert cropped.shape == (inZ, oinY, oinX)
        want = oracle.cached_interpolate(int(method), opx, opy, oinX, oinY, outX, outY, cropped)
        assert_bit_equal(ci.interpolateValues(cropped), want, f"cropped, method {method}")
        # and the crop changes nothing: same values as on the full grid
        full = oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, field)
        assert_bit_equal(want, full, "crop invariance (oracle)")
        assert not ci.createReducedDomain() or ci.reducedDomain()[2:4] == (ominX, ominY)  # "don't set twice"
    # degenerate: all positions in one cell column -> (maxX - minX) >= 1 still holds because of EXTEND; a 1-wide grid does not reduce
    ci = fb.CachedInterpolation("x", "y", Method.BILINEAR, np.zeros(4), np.zeros(4), 1, 1, 2, 2)
    assert not ci.createReducedDomain()


def test_edge_cases(oracle):
    # empty slice (inZ == 0), a single target point, everything outside, ragged level sizes
    ci = fb.CachedInterpolation("x", "y", Method.BILINEAR, [0.5], [0.5], 2, 2, 1, 1)
    out = ci.interpolateValues(np.zeros(0, dtype=np.float32))
    assert out.size == 0
    out = ci.interpolateValues(np.array([1, 2, 3, 4], dtype=np.float32))
    assert out.shape == (1, 1, 1) and out[0, 0, 0] == 2.5
    # size not a multiple of inX*inY: the remainder is ignored like the reference's integer division
    out = ci.interpolateValues(np.array([1, 2, 3, 4, 9, 9], dtype=np.float32))
    assert out.shape == (1, 1, 1)
    ci = fb.CachedInterpolation("x", "y", Method.NEAREST_NEIGHBOR, np.full(7, -999.0), np.full(7, -999.0), 3, 3, 7, 1)
    out = ci.interpolateValues(np.arange(18, dtype=np.float32))
    assert out.shape == (2, 1, 7) and np.all(out.view(np.uint32) == 0x7fc00000)  # canonical quiet NaN fill mask
    # NaN taps poison the result even with zero weight (SURVEY trap 1); inf arithmetic like the reference
    f = np.array([1, np.nan, 3, 4, np.inf, 1, 1, 1], dtype=np.float32)
    px, py = np.array([0.0, 0.25]), np.array([0.0, 0.5])
    ci = fb.CachedInterpolation("x", "y", Method.BILINEAR, px, py, 2, 2, 2, 1)
    got = ci.interpolateValues(f)
    want = oracle.cached_interpolate(1, px, py, 2, 2, 2, 1, f)
    assert_bit_equal(got, want, "NaN/inf taps")
    assert np.isnan(got[0, 0, 0])
    # unknown method / wrong table size
    with pytest.raises(fb.FimexB200Error):
        fb.CachedInterpolation("x", "y", 99, [0.0], [0.0], 2, 2, 1, 1)
    with pytest.raises(fb.FimexB200Error):
        fb.CachedInterpolation("x", "y", Method.FORWARD_MEAN, [0.0], [0.0], 2, 2, 1, 1)


def test_seam_and_border_traps(oracle):
    # SURVEY traps 2, 3, 5: lround ties, half-cell strips, no wrap across the longitude seam
    lon = np.radians(np.arange(1440) * 0.25)
    pts = np.radians(np.array([-0.1, 359.9, 359.75, 0.0, 179.99, -179.9, 360.1]))
    rc, pos = fb.mifi_points2position(pts, lon, fb.LONGITUDE)
    want = oracle.points2position(pts, lon, 1)
    assert np.array_equal(pos.view(np.uint64), want.view(np.uint64))
    assert pos[0] == pytest.approx(-0.4, abs=1e-9)  # probe quoted in SURVEY.md 8a trap 5
    field = np.arange(4 * 1440, dtype=np.float32).reshape(1, 4, 1440)
    py = np.full(pos.size, 1.5)
    for m in (Method.NEAREST_NEIGHBOR, Method.BILINEAR, Method.BICUBIC):
        ci = fb.CachedInterpolation("x", "y", m, pos, py, 1440, 4, pos.size, 1)
        assert_bit_equal(ci.interpolateValues(field), oracle.cached_interpolate(int(m), pos, py, 1440, 4, pos.size, 1, field), f"seam {m}")


# =====================================================================================================
# fused u/v interpolation + rotation (A13, A14) and the matrix (A15)
# =====================================================================================================
@pytest.mark.parametrize("method", [Method.NEAREST_NEIGHBOR, Method.BILINEAR, Method.BICUBIC])
@pytest.mark.parametrize("outX,outY", [(33, 9), (64, 32)])
def test_fused_vector_bit_exact(oracle, method, outX, outY):
    inX, inY, inZ = 50, 40, 6
    rng = np.random.default_rng(int(method) + outX)
    u, px, py = _random_case(21, inX, inY, inZ, outX, outY)
    v = rng.normal(0, 8, u.shape).astype(np.float32)
    phi = rng.uniform(-np.pi, np.pi, outX * outY)
    matrix = np.stack([np.cos(phi), np.sin(phi), -np.sin(phi), phi], axis=1).ravel()
    ui = oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, u)
    vi = oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, v)
    wu, wv = oracle.vector_reproject_by_matrix(matrix, ui, vi, outX, outY, inZ)
    ci = fb.CachedInterpolation("x", "y", method, px, py, inX, inY, outX, outY)
    cvr = fb.CachedVectorReprojection(fb.MIFI_VECTOR_KEEP_SIZE, matrix, outX, outY)
    gu, gv = ci.interpolateVector(u, v, cvr)
    assert_bit_equal(gu, wu.reshape(gu.shape), "fused u")
    assert_bit_equal(gv, wv.reshape(gv.shape), "fused v")
    # un-fused: interpolateValues twice + reprojectValues, as the reference host does (CDMInterpolator.cc:255-276)
    a, b = ci.interpolateValues(u).copy(), ci.interpolateValues(v).copy()
    cvr.reprojectValues(a, b)
    assert_bit_equal(a, gu, "un-fused u")
    assert_bit_equal(b, gv, "un-fused v")
    # no rotation handle: plain interpolation of both
    pu, pv = ci.interpolateVector(u, v, None)
    assert_bit_equal(pu, ui, "no-rotation u")
    assert_bit_equal(pv, vi, "no-rotation v")
    assert np.array_equal(cvr.getMatrix().view(np.uint64), matrix.view(np.uint64))


def test_uninitialised_vector_reprojection_is_identity():
    # CachedVectorReprojection.cc:37-40
    cvr = fb.CachedVectorReprojection()
    u = np.arange(6, dtype=np.float32)
    v = np.arange(6, dtype=np.float32)[::-1].copy()
    cvr.reprojectValues(u, v)
    assert np.array_equal(u, np.arange(6)) and np.array_equal(v, np.arange(6)[::-1])


@pytest.mark.parametrize("target,axes,types", [
    (ROTPOLE, (np.linspace(-10, 10, 41), np.linspace(-8, 8, 33)), (1, 2)),
    (STERE, (np.linspace(-2e6, 2e6, 41), np.linspace(-2e6, 2e6, 33)), (0, 0)),
    (LCC, (np.linspace(-1e6, 1e6, 41), np.linspace(-2e6, 2e6, 33)), (0, 0)),
])
def test_vector_matrix_against_oracle(oracle, target, axes, types):
    rc, want = oracle.vector_matrix(SRC_LL, target, axes[0], axes[1], types[0], types[1])
    assert rc == 1
    rc, got = fb.mifi_get_vector_reproject_matrix(SRC_LL, target, axes[0], axes[1], types[0], types[1])
    assert rc == fb.MIFI_OK
    assert np.abs(got[3::4] - want[3::4]).max() <= 1e-7 * DEG
    assert np.abs(got[0::4]**2 + got[1::4]**2 - 1).max() < 1e-14
    cvr = fb.CachedVectorReprojection.fromProjection(0, SRC_LL, target, axes[0], axes[1], types[0], types[1])
    assert np.array_equal(cvr.getMatrix().view(np.uint64), got.view(np.uint64))


# =====================================================================================================
# forward interpolation (A16, A17)
# =====================================================================================================
@pytest.mark.parametrize("method", [Method.FORWARD_MEAN, Method.FORWARD_MAX, Method.FORWARD_SUM, Method.FORWARD_MIN, Method.FORWARD_MEDIAN,
                                    Method.FORWARD_UNDEF_MEAN, Method.FORWARD_UNDEF_MAX, Method.FORWARD_UNDEF_SUM, Method.FORWARD_UNDEF_MIN])
@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("inZ", [3, 11])  # 11: two passes of four levels per thread + a tail of three single levels
def test_forward_bit_exact(oracle, method, dense, inZ):
    rng = np.random.default_rng(int(method) * 2 + dense)
    inX, inY = 211, 57
    outX, outY = (19, 13) if dense else (160, 90)  # dense: ~50 points per cell; sparse: many empty cells
    n = inX * inY
    px = rng.uniform(-2, outX + 1, n)
    py = rng.uniform(-2, outY + 1, n)
    px[:50] = np.round(px[:50] * 2) / 2  # x.5 ties: round() goes away from zero
    px[50:60] = np.nan
    py[60:70] = -999.0
    data = rng.normal(10, 5, (inZ, inY, inX)).astype(np.float32)
    data[rng.random(data.shape) < 0.05] = np.nan
    want = oracle.forward_interpolate(int(method), px, py, inX, inY, outX, outY, data)
    cfi = fb.CachedForwardInterpolation("x", "y", method, px, py, inX, inY, outX, outY)
    got = cfi.interpolateValues(data)
    if method in (Method.FORWARD_UNDEF_MAX, Method.FORWARD_UNDEF_MIN):
        # NaN payload/sign may differ where the aggregate is NaN; compare the mask and the numbers
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
    else:
        assert_bit_equal(got, want, f"forward method {method} dense={dense}")
    assert np.isnan(got).any() or dense


def test_forward_from_coordinates_against_oracle(oracle):
    # the table-producing part of changeProjectionByForwardInterpolation + the scatter, swath-like input
    rng = np.random.default_rng(3)
    ny, nx = 120, 80
    lat = 60 + np.linspace(0, 4, ny)[:, None] + np.linspace(0, 0.4, nx)[None, :]
    lon = 10 + np.linspace(0, 1, ny)[:, None] + np.linspace(0, 6, nx)[None, :]
    lat += rng.normal(0, 1e-3, lat.shape)
    ox = np.arange(-200e3, 200e3 + 1, 5e3)
    oy = np.arange(-350e3, 150e3 + 1, 5e3)
    data = rng.normal(0, 1, (1, ny, nx)).astype(np.float32)
    data[rng.random(data.shape) < 0.02] = np.nan
    # oracle pipeline (CDMInterpolator.cc:1265-1317)
    rc, x, y = oracle.project_values(WGS84, LCC, np.radians(lon.ravel()), np.radians(lat.ravel()))
    assert rc == 1
    x = oracle.points2position(x, ox, 0)
    y = oracle.points2position(y, oy, 0)
    for m in (Method.FORWARD_MEAN, Method.FORWARD_MAX):
        want = oracle.forward_interpolate(int(m), x, y, nx, ny, ox.size, oy.size, data)
        cfi = fb.CachedForwardInterpolation.fromCoordinates(m, LCC, ox, oy, False, False, lon.ravel(), lat.ravel(), nx, ny)
        got = cfi.interpolateValues(data)
        gx, gy = cfi.points()
        assert np.abs(gx - x).max() < 1e-6 and np.abs(gy - y).max() < 1e-6  # cell units; 1e-9 deg ~ 2e-8 cells here
        flips = (np.round(gx) != np.round(x)).sum() + (np.round(gy) != np.round(y)).sum()
        if flips == 0:
            assert_bit_equal(got, want, f"forward from coordinates {m}")
        else:  # a point within 1e-6 of a cell edge moved: bounded, reported
            assert flips <= 3 and (got.view(np.uint32) != want.view(np.uint32)).sum() <= 2 * flips


# =====================================================================================================
# table production on the device (A9, A10, A18, A19)
# =====================================================================================================
def _config2_like(n=200):
    lon = np.arange(1440) * 0.25
    lat = 90 - np.arange(721) * 0.25
    ax = (np.arange(n) - (n - 1) / 2) * (45.0 / n)
    return lon, lat, ax


@pytest.mark.parametrize("method", [Method.NEAREST_NEIGHBOR, Method.BILINEAR, Method.BICUBIC])
def test_from_projection_against_oracle(oracle, method):
    lon, lat, ax = _config2_like(160)
    # oracle pipeline: CDMInterpolator.cc:1446-1476 + createReducedDomain
    rc, x, y = oracle.project_axes(ROTPOLE, SRC_LL, np.radians(ax), np.radians(ax))
    assert rc == 1
    px = oracle.points2position(x, np.radians(lon), 1)
    py = oracle.points2position(y, np.radians(lat), 2)
    red, opx, opy, inX, inY, x0, y0 = oracle.reduced_domain(px, py, 1440, 721)
    assert red
    ci = fb.CachedInterpolation.fromProjection(method, ROTPOLE, ax, ax, True, True, SRC_LL, lon, lat, True)
    assert ci.createReducedDomain()
    assert (ci.getInX(), ci.getInY()) == (inX, inY) and ci.reducedDomain()[2:4] == (x0, y0)
    gx, gy = ci.points()
    # 1e-9 degree on a 0.25-degree grid is 4e-9 cells
    assert np.abs(gx - opx).max() <= 4e-9 and np.abs(gy - opy).max() <= 4e-9
    rng = np.random.default_rng(2)
    field = rng.normal(250, 30, (4, inY, inX)).astype(np.float32)
    got = ci.interpolateValues(field)
    # with the device's own positions the gather is bit-exact against the oracle ...
    assert_bit_equal(got, oracle.cached_interpolate(int(method), gx, gy, inX, inY, ax.size, ax.size, field), "own positions")
    # ... and against the oracle's positions only points within ~1e-9 cells of a cell/half-cell boundary may differ
    want = oracle.cached_interpolate(int(method), opx, opy, inX, inY, ax.size, ax.size, field)
    if method == Method.NEAREST_NEIGHBOR:
        flips = (got[0] != want[0]).sum()
        assert flips <= 2, flips
    else:
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
        assert np.nanmax(rel) <= 1e-5


def test_coord_nearestneighbor_against_oracle(oracle):
    lon = np.arange(0, 360, 1.5)
    lat = 90 - np.arange(0, 180.1, 1.5)
    ax = (np.arange(60) - 29.5) * 0.7
    lon2d, lat2d = oracle.lonlat_to_matrix(np.radians(lon), np.radians(lat))
    rc, tx, ty = oracle.project_axes(ROTPOLE, WGS84, np.radians(ax), np.radians(ax))
    assert rc == 1
    wx, wy, ties = oracle.coordnn(tx, ty, lon2d, lat2d, lon.size, lat.size)
    ci = fb.CachedInterpolation.fromCoordinates(Method.COORD_NN, ROTPOLE, ax, ax, True, True, np.degrees(lon2d), np.degrees(lat2d), lon.size,
                                                lat.size)
    gx, gy = ci.points()
    differ = ((gx != wx) | (gy != wy)).sum()
    assert differ <= max(2, ties), (differ, ties)  # only (near-)equidistant candidates may resolve differently
    assert (gx >= 0).all() and (gy >= 0).all()
    field = np.random.default_rng(1).normal(0, 1, (2, lat.size, lon.size)).astype(np.float32)
    got = ci.interpolateValues(field)
    assert_bit_equal(got, oracle.cached_interpolate(3, gx, gy, lon.size, lat.size, ax.size, ax.size, field), "coord_nn gather")
    # a target far from every source point gets (-1,-1) -> NaN (LL_POINT default, CDMInterpolator.cc:1134)
    ci2 = fb.CachedInterpolation.fromCoordinates(Method.COORD_NN, SRC_LL, [10.0, 200.0], [45.0], True, True, [10.0, 10.5, 10.0, 10.5],
                                                 [45.0, 45.0, 45.5, 45.5], 2, 2)
    gx, gy = ci2.points()
    assert gx[0] == 0 and gy[0] == 0 and gx[1] == -1 and gy[1] == -1


def test_latlon_template_against_oracle(oracle):
    """changeProjectionByProjectionParametersToLatLonTemplate (CDMInterpolator.cc:1755-1823): a curvilinear lon/lat template
    (and a 1-D point list) as the target: project_values + points2position on the device, matrix from points"""
    lon, lat, _ = _config2_like(40)
    rng = np.random.default_rng(4)
    ny, nx = 30, 50
    jj, ii = np.meshgrid(np.arange(nx), np.arange(ny))
    tlon = (-20 + 0.6 * jj + 0.1 * ii + 0.3 * np.sin(ii / 5.0)).ravel()
    tlat = (50 + 0.4 * ii - 0.05 * jj + 0.2 * np.cos(jj / 7.0)).ravel()
    field = rng.normal(250, 30, (3, lat.size, lon.size)).astype(np.float32)
    for method in (Method.BILINEAR, Method.NEAREST_NEIGHBOR, Method.BICUBIC):
        ci = fb.CachedInterpolation.fromTemplate(method, WGS84, tlon, tlat, nx, ny, SRC_LL, lon, lat, True)
        gx, gy = ci.points()
        rc, x, y = oracle.project_values(WGS84, SRC_LL, np.radians(tlon), np.radians(tlat))
        assert rc == 1
        wy = oracle.points2position(y, np.radians(lat), 2)
        wx = oracle.points2position(x, np.radians(lon), 1)
        assert np.abs(gx - wx).max() < 1e-8 and np.abs(gy - wy).max() < 1e-8  # 1e-9 degree = 4e-9 cells here
        got = ci.interpolateValues(field)
        assert_bit_equal(got, oracle.cached_interpolate(int(method), gx, gy, lon.size, lat.size, nx, ny, field), f"template {method}",
                         nan_payload=(method == Method.NEAREST_NEIGHBOR))
        assert ci.createReducedDomain()
    # cross-section style: a list of points (outY = 1)
    ci = fb.CachedInterpolation.fromTemplate(Method.BILINEAR, WGS84, tlon[:77], tlat[:77], 77, 1, SRC_LL, lon, lat, True)
    assert ci.interpolateValues(field).shape == (3, 1, 77)
    # rotation matrix for the template points: source rotated pole -> geographic, angles against the oracle
    rlon = rng.uniform(-10, 10, 200)
    rlat = rng.uniform(-10, 10, 200)
    cvr = fb.CachedVectorReprojection.fromPoints(fb.MIFI_VECTOR_KEEP_SIZE, ROTPOLE, WGS84, 0, rlon, rlat)
    rc, m = oracle.vector_matrix_points(ROTPOLE, WGS84, 0, np.radians(rlon), np.radians(rlat))
    assert rc == 1
    g = cvr.getMatrix()
    assert np.abs(g[0::4] - m[0::4]).max() < 1e-7 and np.abs(g[1::4] - m[1::4]).max() < 1e-7


def test_coord_kdtree_against_oracle(oracle):
    """coord_kdtree (MIFI_INTERPOL_COORD_NN_KD, CDMInterpolator.cc:991-1062, 1406-1408): nearest source point by squared chord
    distance inside the distance of interest, else (-1000, -1000) -> NaN"""
    lon = np.arange(0, 360, 1.5)
    lat = 90 - np.arange(0, 180.1, 1.5)
    ax = (np.arange(60) - 29.5) * 0.7
    lon2d, lat2d = oracle.lonlat_to_matrix(np.radians(lon), np.radians(lat))
    rc, tx, ty = oracle.project_axes(ROTPOLE, WGS84, np.radians(ax), np.radians(ax))
    assert rc == 1
    for max_dist in (0.0, 60e3, 2000e3):
        # 0: derived from the axes like getMaxDistanceOfInterest -- degrees times the earth radius, as the reference does
        dist = max_dist if max_dist > 0 else oracle.max_distance_of_interest(ax, ax, False)
        wx, wy, ties = oracle.coordkd(tx, ty, lon2d, lat2d, lon.size, lat.size, dist)
        ci = fb.CachedInterpolation.fromCoordinates(Method.COORD_NN_KD, ROTPOLE, ax, ax, True, True, np.degrees(lon2d), np.degrees(lat2d),
                                                    lon.size, lat.size, maxDistance=max_dist)
        gx, gy = ci.points()
        differ = int(((gx != wx) | (gy != wy)).sum())
        assert differ <= max(2, ties), (max_dist, differ, ties)  # only (near-)equidistant candidates may resolve differently
        if max_dist == 60e3:
            assert (wx == -1000).any() and (wx >= 0).any()  # 1.5-degree cells: most targets are further than 60 km from a centre
        field = np.random.default_rng(1).normal(0, 1, (2, lat.size, lon.size)).astype(np.float32)
        assert_bit_equal(ci.interpolateValues(field), oracle.cached_interpolate(4, gx, gy, lon.size, lat.size, ax.size, ax.size, field),
                         "coord_kdtree gather", nan_payload=True)
    with pytest.raises(fb.FimexB200Error):
        fb.CachedInterpolation.fromCoordinates(Method.BILINEAR, ROTPOLE, ax, ax, True, True, np.degrees(lon2d), np.degrees(lat2d), lon.size,
                                               lat.size)


def test_interpolator_config1_hirlam12_like(oracle):
    # BASELINE config 1: test/hirlam12.nc geometry (Xc = 5.0..6.6 step 0.1, Yc = 61.5..62.6, 2 levels x 2 times) bilinear
    # to a 0.5-degree lat/long grid via the option strings of the fimex CLI (src/binSrc/fimex.cc:1008-1034)
    xc = 5.0 + 0.1 * np.arange(17)
    yc = 61.5 + 0.1 * np.arange(12)
    rng = np.random.default_rng(12)
    fill = np.float32(9.96921e+36)
    data = rng.normal(280, 5, (2, 2, 12, 17)).astype(np.float32)
    data[0, 1, 5, 5] = fill  # (lon 5.5, lat 62.0): a tap of the target point (5.5, 62.0)
    xw = rng.normal(0, 5, data.shape).astype(np.float32)
    yw = rng.normal(0, 5, data.shape).astype(np.float32)
    proj_out = "+proj=latlong +a=6371000 +e=0 +no_defs"
    ip = fb.Interpolator(SRC_LL, xc, yc, True, has_xy_vectors=True)
    ip.changeProjection("bilinear", proj_out, "5,5.5,6,6.5", "61.5,62,62.5", "degree", "degree")
    out = ip.getDataSlice(data, bad_value=fill)
    assert out.shape == (2, 2, 3, 4)
    # oracle: same pipeline on the CPU
    rc, x, y = oracle.project_axes(proj_out, SRC_LL, np.radians([5, 5.5, 6, 6.5]), np.radians([61.5, 62, 62.5]))
    px = oracle.points2position(x, np.radians(xc), 1)
    py = oracle.points2position(y, np.radians(yc), 2)
    red, px, py, inX, inY, x0, y0 = oracle.reduced_domain(px, py, 17, 12)
    crop = data[..., y0:y0 + inY, x0:x0 + inX]
    want = oracle.cached_interpolate(1, px, py, inX, inY, 4, 3, oracle.bad2nan(crop, fill))
    want = np.where(np.isnan(want), fill, want).reshape(out.shape)
    assert_bit_equal(out, want, "config 1 scalar")
    assert (out == fill).sum() >= 1
    # x_wind / y_wind: rotated with the lat/long-target bearing branch of the matrix (interpolation.c:366,408)
    ux = ip.getDataSlice(xw, counterpart=yw, direction="x")
    uy = ip.getDataSlice(yw, counterpart=xw, direction="y")
    rc, m = oracle.vector_matrix(SRC_LL, proj_out, [5, 5.5, 6, 6.5], [61.5, 62, 62.5], 1, 2)
    ui = oracle.cached_interpolate(1, px, py, inX, inY, 4, 3, xw[..., y0:y0 + inY, x0:x0 + inX])
    vi = oracle.cached_interpolate(1, px, py, inX, inY, 4, 3, yw[..., y0:y0 + inY, x0:x0 + inX])
    wu, wv = oracle.vector_reproject_by_matrix(m, ui, vi, 4, 3, 4)
    assert np.allclose(ux.ravel(), wu.ravel(), rtol=1e-5, atol=1e-6) and np.allclose(uy.ravel(), wv.ravel(), rtol=1e-5, atol=1e-6)


# =====================================================================================================
# A1/A2: the whole slice body of getDataSlice in one call -- fill -> NaN while loading, NaN -> fill + round + cast
# while storing (CDMInterpolator.cc:115-124, 250-285)
# =====================================================================================================
# (input type, variable type, fill value; None = the type's NetCDF default as CDM::getFillValue returns it)
_TYPED = [(np.float32, np.float32, None), (np.int16, np.int16, None), (np.float32, np.int16, -32767.0), (np.int16, np.float32, -32767.0),
          (np.float64, np.float64, None), (np.int32, np.int32, None), (np.uint8, np.uint8, None), (np.int8, np.int8, None),
          (np.uint16, np.uint16, None), (np.uint32, np.uint32, 4000000000.0), (np.int64, np.int64, -9999.0), (np.uint64, np.uint64, 12345.0),
          (np.float32, np.float32, np.nan)]


def _typed_field(rng, shape, dtype, fill):
    dt = np.dtype(dtype)
    if dt.kind == "f":
        a = rng.normal(250, 30, shape).astype(dt)
    else:
        info = np.iinfo(dt)
        lo, hi = max(info.min, -30000), min(info.max, 30000)
        a = rng.integers(lo, hi, shape, endpoint=True).astype(dt)
    a[rng.random(shape) < 0.03] = dt.type(fill)
    return a


@pytest.mark.parametrize("method", [Method.BILINEAR, Method.NEAREST_NEIGHBOR, Method.BICUBIC, Method.COORD_NN])
@pytest.mark.parametrize("in_dtype,out_dtype,fill", _TYPED)
def test_get_data_slice_typed(oracle, method, in_dtype, out_dtype, fill):
    inX, inY, inZ, outX, outY = 60, 50, 11, 200, 90
    px, py = _smooth_positions(inX, inY, outX, outY, 17.0, 5.0, 3)
    rng = np.random.default_rng(int(method) * 100 + np.dtype(in_dtype).itemsize)
    if fill is None:
        fill = fb.default_fill_value(in_dtype)
    data = _typed_field(rng, (inZ, inY, inX), in_dtype, fill)
    if np.isnan(fill):
        data[0, 0, 0] = -0.0  # a NaN fill value switches the input pass off; the output pass still turns -0 into +0
    ci = fb.CachedInterpolation("x", "y", method, px, py, inX, inY, outX, outY)
    got = ci.getDataSlice(data, fill, out_dtype)
    assert got.dtype == np.dtype(out_dtype) and got.shape == (inZ, outY, outX)
    interp = oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, oracle.as_float(data, fill))
    want = oracle.from_float(interp, fill, out_dtype).reshape(got.shape)
    if np.dtype(out_dtype).kind == "f":
        bits = np.uint32 if got.itemsize == 4 else np.uint64
        same = (got.view(bits) == want.view(bits)) | (np.isnan(got) & np.isnan(want))  # NaN only when the fill value is NaN
        assert same.all()
    else:
        assert np.array_equal(got, want)
    if np.isnan(fill):
        assert np.isnan(want).any()
    else:
        assert (want == np.dtype(out_dtype).type(fill)).any()  # the fill value came through both adapters


@pytest.mark.parametrize("method", [Method.BILINEAR, Method.BICUBIC, Method.FORWARD_MEAN])
def test_get_data_slice_device_and_unfused_paths(oracle, method, monkeypatch):
    import torch
    inX, inY, inZ, outX, outY = 60, 50, 70, 203, 77  # outX not a multiple of 4; two level chunks
    rng = np.random.default_rng(5)
    fill = -32767.0
    data = _typed_field(rng, (inZ, inY, inX), np.int16, fill)
    if method == Method.FORWARD_MEAN:
        px = rng.uniform(-1, outX, inX * inY)
        py = rng.uniform(-1, outY, inX * inY)
        ci = fb.CachedForwardInterpolation("x", "y", method, px, py, inX, inY, outX, outY)
        interp = oracle.forward_interpolate(int(method), px, py, inX, inY, outX, outY, oracle.as_float(data, fill))
    else:
        px, py = _smooth_positions(inX, inY, outX, outY, -33.0, 6.5, 3)
        ci = fb.CachedInterpolation("x", "y", method, px, py, inX, inY, outX, outY)
        interp = oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, oracle.as_float(data, fill))
    want = oracle.from_float(interp, fill, np.int16).reshape(inZ, outY, outX)
    host = ci.getDataSlice(data, fill)
    assert np.array_equal(host, want)
    dev = ci.getDataSlice(torch.from_numpy(data).cuda(), fill)
    assert dev.dtype == torch.int16 and np.array_equal(dev.cpu().numpy(), want)
    # float in / float out with a fill value: in-kernel compare + in-kernel NaN -> fill
    fdata = oracle.as_float(data, np.nan)
    f32fill = fb.default_fill_value(np.float32)
    fdata[fdata == fill] = f32fill
    gotf = ci.getDataSlice(torch.from_numpy(fdata).cuda(), f32fill).cpu().numpy()
    wantf = oracle.from_float(interp, f32fill, np.float32).reshape(gotf.shape)
    assert np.array_equal(gotf.view(np.uint32), wantf.view(np.uint32))
    if method != Method.FORWARD_MEAN:  # the same through the direct (unstaged) kernels + conversion passes
        monkeypatch.setenv("FIMEX_B200_DIRECT_GATHER", "1")
        cd = fb.CachedInterpolation("x", "y", method, px, py, inX, inY, outX, outY)
        monkeypatch.delenv("FIMEX_B200_DIRECT_GATHER")
        assert np.array_equal(cd.getDataSlice(data, fill), want)
        assert np.array_equal(cd.getDataSlice(fdata, f32fill).view(np.uint32), wantf.view(np.uint32))


@pytest.mark.parametrize("method", [Method.BILINEAR, Method.BICUBIC, Method.NEAREST_NEIGHBOR])
@pytest.mark.parametrize("dtype", [np.float32, np.int16])
def test_get_vector_slice_typed(oracle, method, dtype):
    inX, inY, inZ, outX, outY = 60, 50, 9, 200, 90
    px, py = _smooth_positions(inX, inY, outX, outY, 17.0, 5.0, 3)
    rng = np.random.default_rng(31)
    fu, fv = fb.default_fill_value(dtype), (-9999.0 if np.dtype(dtype).kind == "f" else -9999)
    u = _typed_field(rng, (inZ, inY, inX), dtype, fu)
    v = _typed_field(rng, (inZ, inY, inX), dtype, fv)
    phi = rng.uniform(-np.pi, np.pi, outX * outY)
    matrix = np.stack([np.cos(phi), np.sin(phi), -np.sin(phi), phi], axis=1).ravel()
    ci = fb.CachedInterpolation("x", "y", method, px, py, inX, inY, outX, outY)
    cvr = fb.CachedVectorReprojection(fb.MIFI_VECTOR_KEEP_SIZE, matrix, outX, outY)
    gu, gv = ci.getVectorSlice(u, v, fu, fv, cvr)
    ui = oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, oracle.as_float(u, fu))
    vi = oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, oracle.as_float(v, fv))
    wu, wv = oracle.vector_reproject_by_matrix(matrix, ui, vi, outX, outY, inZ)
    wu = oracle.from_float(wu, fu, dtype).reshape(gu.shape)
    wv = oracle.from_float(wv, fv, dtype).reshape(gv.shape)
    if np.dtype(dtype).kind == "f":
        assert np.array_equal(gu.view(np.uint32), wu.view(np.uint32)) and np.array_equal(gv.view(np.uint32), wv.view(np.uint32))
    else:
        assert np.array_equal(gu, wu) and np.array_equal(gv, wv)
    assert (wu == np.dtype(dtype).type(fu)).any() and (wv == np.dtype(dtype).type(fv)).any()


# =====================================================================================================
# full BASELINE size (config 2): properties that need no oracle run over 5e8 values
# =====================================================================================================
def test_full_size_bilinear_properties(oracle):
    import torch
    lon = np.arange(1440) * 0.25
    lat = 90 - np.arange(721) * 0.25
    ax = (np.arange(2000) - 999.5) * 0.0225
    ci = fb.CachedInterpolation.fromProjection(Method.BILINEAR, ROTPOLE, ax, ax, True, True, SRC_LL, lon, lat, True)
    assert ci.createReducedDomain()
    inX, inY, nz = ci.getInX(), ci.getInY(), 137
    g = torch.Generator(device="cuda").manual_seed(20261018)
    a = torch.randn((nz, inY, inX), generator=g, device="cuda", dtype=torch.float32)
    b = torch.randn((nz, inY, inX), generator=g, device="cuda", dtype=torch.float32)
    oa = ci.interpolateValues(a)
    assert oa.shape == (nz, 2000, 2000) and not torch.isnan(oa).any()
    # (1) a plane a*x + b*y + c is reproduced (bilinear is exact for it) to fp32 rounding
    xs = torch.arange(inX, device="cuda", dtype=torch.float32)[None, None, :]
    ys = torch.arange(inY, device="cuda", dtype=torch.float32)[None, :, None]
    plane = (0.5 * xs + 0.25 * ys + 3.0).expand(2, inY, inX).contiguous()
    op = ci.interpolateValues(plane)
    px, py = ci.points()
    want = torch.from_numpy((0.5 * px + 0.25 * py + 3.0).astype(np.float32)).cuda().view(2000, 2000)
    interior = torch.from_numpy((px >= 0) & (px <= inX - 1) & (py >= 0) & (py <= inY - 1)).cuda().view(2000, 2000)
    assert interior.float().mean().item() > 0.99  # only the half cells outside the 0/360 seam fall back to nearest (trap 5)
    assert ((op[0] - want).abs() * interior).max().item() <= 2e-5 * want.abs().max().item()
    # (2) linearity: I(a + b) == I(a) + I(b) within fp32 rounding of the sums
    ob = ci.interpolateValues(b)
    oab = ci.interpolateValues(a + b)
    assert (oab - (oa + ob)).abs().max().item() <= 1e-5 * 8
    # (3) a random sample of target columns against the oracle, all 137 levels, bit for bit
    rng = np.random.default_rng(0)
    sample = rng.integers(0, 2000 * 2000, 2000)
    host = a.cpu().numpy()
    want = oracle.cached_interpolate(1, px[sample], py[sample], inX, inY, sample.size, 1, host)
    got = oa.view(nz, -1)[:, torch.from_numpy(sample).cuda()].cpu().numpy().reshape(want.shape)
    assert_bit_equal(got, want, "full-size sample vs oracle")
    # (4) checksum of checksums: per-level sums of the output equal the sums of per-row sums (layout sanity)
    assert torch.allclose(oa.sum(dim=(1, 2), dtype=torch.float64), oa.sum(dim=2, dtype=torch.float64).sum(dim=1))


# =====================================================================================================
# 8f rank 3: fill2d / creepfill2d pre/post-processes on the device (anti-diagonal wavefronts of the reference's sweeps)
# =====================================================================================================
def _holey(rng, shape, frac=0.1):
    ny, nx = shape[-2:]
    yy, xx = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    f = (280 + 10 * np.sin(xx / 7.0) * np.cos(yy / 5.0) + rng.normal(0, 0.5, shape)).astype(np.float32)
    f[rng.random(shape) < frac] = np.nan
    if ny > 12 and nx > 20:
        f[..., 3:9, 4:15] = np.nan
        f[..., 0, :5] = np.nan
        f[..., -1, -3:] = np.nan
        f[..., 5:8, 0] = np.nan
    return f


@pytest.mark.parametrize("shape", [(3, 20, 31), (2, 77, 130), (1, 2, 2), (2, 2, 9), (2, 9, 2), (1, 3, 3), (1, 600, 3)])
def test_fill2d_creepfill2d_bit_exact(oracle, shape):
    import torch
    rng = np.random.default_rng(sum(shape))
    f = _holey(rng, shape, 0.3 if min(shape[-2:]) < 10 else 0.1)
    for args in ((0.01, 1.6, 100), (1e-6, 1.0, 7), (0.5, 1.9, 3), (0.01, 1.6, 0)):
        want, counts = oracle.fill2d(f, *args)
        dev = fb.fill2d_device(torch.from_numpy(f.copy()).cuda(), *args).cpu().numpy()
        assert np.array_equal(dev.view(np.uint32), want.view(np.uint32)), ("fill2d", args)
        rc, one, n = fb.mifi_fill2d_f(f[0], *args)  # the reference's own host-pointer prototype
        assert rc == fb.MIFI_OK and n == counts[0] and np.array_equal(one.view(np.uint32), want[0].view(np.uint32))
    for repeat, weight, dv in ((20, 2, None), (3, 1, None), (5, 2, 0.0), (1, 3, -7.5), (0, 2, None)):
        want, counts = oracle.creepfill2d(f, repeat, weight, dv)
        dev = fb.creepfill2d_device(torch.from_numpy(f.copy()).cuda(), repeat, weight, dv).cpu().numpy()
        assert np.array_equal(dev.view(np.uint32), want.view(np.uint32)), ("creepfill2d", repeat, weight, dv)
        rc, one, n = fb.mifi_creepfill2d_f(f[0], repeat, weight, dv)
        assert rc == fb.MIFI_OK and n == counts[0] and np.array_equal(one.view(np.uint32), want[0].view(np.uint32))
    # nothing to fill / nothing defined: untouched
    full = np.ones(shape, dtype=np.float32)
    assert np.array_equal(fb.fill2d_device(torch.from_numpy(full.copy()).cuda(), 0.01, 1.6, 100).cpu().numpy(), full)
    allnan = torch.full(shape, float("nan"), device="cuda")
    assert torch.isnan(fb.creepfill2d_device(allnan, 20, 2)).all()


def test_fill2d_as_preprocess_at_source_size(oracle):
    """the cropped ERA5 footprint of config 2 (1440 x 202), a handful of levels: timing + parity of one level"""
    import torch
    rng = np.random.default_rng(3)
    f = _holey(rng, (8, 202, 1440), 0.05)
    d = torch.from_numpy(f.copy()).cuda()
    fb.fill2d_device(d, 0.01, 1.6, 100)
    torch.cuda.synchronize()
    want, _ = oracle.fill2d(f[:1], 0.01, 1.6, 100)
    assert np.array_equal(d[:1].cpu().numpy().view(np.uint32), want.view(np.uint32))
    d2 = torch.from_numpy(f.copy()).cuda()
    fb.creepfill2d_device(d2, 20, 2)
    want, _ = oracle.creepfill2d(f[:1], 20, 2)
    assert np.array_equal(d2[:1].cpu().numpy().view(np.uint32), want.view(np.uint32))


def test_concurrent_calls_on_one_handle(oracle):
    """interpolateValues / getDataSlice / reprojectValues are const in the reference and are called concurrently from the
    writer's OpenMP tasks on ONE object (src/NetCDF_CDMWriter.cc:749-753): eight host threads share a handle here (ctypes
    releases the GIL during the call), each with its own data, host-buffer path (per-call streams and scratch)."""
    import threading
    inX, inY, inZ, outX, outY = 60, 50, 40, 203, 160
    px, py = _smooth_positions(inX, inY, outX, outY, 17.0, 5.0, 3)
    rng = np.random.default_rng(77)
    fields = [rng.normal(250, 30, (inZ, inY, inX)).astype(np.float32) for _ in range(8)]
    for method in (Method.BILINEAR, Method.BICUBIC):
        ci = fb.CachedInterpolation("x", "y", method, px, py, inX, inY, outX, outY)
        want = [oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, f) for f in fields]
        got = [None] * len(fields)
        errors = []

        def work(i):
            try:
                for _ in range(3):
                    got[i] = ci.interpolateValues(fields[i]) if i % 2 == 0 else ci.getDataSlice(fields[i], np.nan)
            except Exception as e:  # pragma: no cover
                errors.append(e)

        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(fields))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors
        for i in range(len(fields)):
            assert_bit_equal(got[i], want[i], f"thread {i}, method {method}")


def test_get_data_slice_with_pre_and_postprocess(oracle):
    """--interpolate.preprocess / postprocess inside the slice call (CDMInterpolator.cc:254-256, 284; parseProcess,
    src/binSrc/fimex.cc:644-671 -- including its one-character weight: "2" is the weight 50)"""
    inX, inY, inZ, outX, outY = 60, 50, 5, 120, 70
    px, py = _smooth_positions(inX, inY, outX, outY, 17.0, 2.2, 3)  # zoom 2.2: part of the target hangs over the source grid
    rng = np.random.default_rng(2)
    fill = -32767.0
    data = _typed_field(rng, (inZ, inY, inX), np.int16, fill)
    data[:, 10:20, 15:30] = -32767
    ci = fb.CachedInterpolation("x", "y", Method.BILINEAR, px, py, inX, inY, outX, outY)
    ci.addPreprocess(" fill2d(0.01,1.6,100)")
    ci.addPreprocess("creepfill2d(3,2)")
    ci.addPostprocess("creepfill2d(2,1,-5.5)")
    got = ci.getDataSlice(data, fill)
    a = oracle.as_float(data, fill)
    a, _ = oracle.fill2d(a, np.float32(0.01), np.float32(1.6), 100)
    a, _ = oracle.creepfill2d(a, 3, ord("2"))
    o = oracle.cached_interpolate(1, px, py, inX, inY, outX, outY, a)
    assert np.isnan(o).any()  # outside the source grid: filled by the postprocess
    o, _ = oracle.creepfill2d(o.reshape(inZ, outY, outX), 2, ord("1"), -5.5)
    want = oracle.from_float(o, fill, np.int16).reshape(got.shape)
    assert np.array_equal(got, want)
    assert not (got == -32767).any()
    for bad in ("fill2d(0.01,1.6)", "creepfill2d(1)", "smooth(3)", "fill2d(a,b,c)"):
        with pytest.raises(fb.FimexB200Error):
            ci.addPreprocess(bad)


def test_pinned_host_buffers_are_recycled():
    """fb200_host_alloc / fb200_host_free: freed page-locked buffers are handed out again (locking pages costs about as much
    as copying them; the reference allocates a new output array per interpolateValues call)"""
    import ctypes as C
    lib = fb.load()
    a = lib.fb200_host_alloc(50 << 20)
    assert a
    C.memset(a, 1, 50 << 20)
    lib.fb200_host_free(a)
    b = lib.fb200_host_alloc(49 << 20)  # same 2 MB size class or slightly smaller: the cached block comes back
    assert b == a
    c = lib.fb200_host_alloc(50 << 20)  # a second one while the first is in use: a different block
    assert c and c != b
    lib.fb200_host_free(b)
    lib.fb200_host_free(c)
    lib.fb200_host_trim()
    d = lib.fb200_host_alloc(1)
    assert d
    lib.fb200_host_free(d)
    lib.fb200_host_trim()
    # a pointer that did not come from fb200_host_alloc is left alone (and reported), not handed to cudaFreeHost
    foreign = np.zeros(1024, dtype=np.float32)
    lib.fb200_host_free(foreign.ctypes.data_as(C.c_void_p))
    assert b"not allocated by fb200_host_alloc" in lib.fb200_last_error()
    assert foreign.sum() == 0
    lib.fb200_host_free(None)


# ---- test_interpolator_vector_backforth (test/testInterpolator.cc:398-472) ----------------------------------------
# The reference's one quantitative end-to-end check of the vector path: a constant unit wind pointing north (or east) on
# a global lat/lon grid is regridded (nearest neighbour, vectors rotated) to a projection and back to lat/lon; whatever
# is not NaN must again be (0,1) / (1,0) within a per-projection delta.  test/data/north.nc / east.nc are NetCDF-4 files
# this image cannot read; their content is the constant field synthesised here.  The utm row is outside the four
# projections of the path and is left out; proj strings, axes and deltas are the reference's.
_R = "6371000"
_BACKFORTH = [
    ("+proj=stere +lat_0=90 +lon_0=-32 +lat_ts=60 +ellps=sphere +R=" + _R, "-30000000,-29950000,...,30000000",
     "-30000000,-29950000,...,30000000", "m", "-180,-179,...,179", "55,56,...,87", 8e-2),
    ("+proj=stere +lat_0=90 +lon_0=0 +lat_ts=90 +ellps=sphere +R=" + _R, "-30000000,-29950000,...,30000000",
     "-30000000,-29950000,...,30000000", "m", "-180,-179,...,179", "55,56,...,87", 8e-2),
    ("+proj=stere +lat_0=-90 +lon_0=0 +lat_ts=-90 +ellps=sphere +R=" + _R, "-30000000,-29950000,...,30000000",
     "-30000000,-29950000,...,30000000", "m", "0,1,359", "-55,-56,...,-87", 8e-2),
    ("+proj=lcc +lat_0=63 +lon_0=15 +lat_1=63 +lat_2=63 +no_defs +R=6.371e+06", "-922000,-902000,...,922000",
     "-1130000,-1110000,...,1230000", "m", "-30,-29,...,40", "50,51,...,85", 1e-2),
    ("+proj=ob_tran +o_proj=longlat +lon_0=-40 +o_lat_p=22 +R=6.371e+06 +no_defs", "16.5,16.6,...,24.2", "-3.8,-3.7,...,14.9",
     "degree", "-30,-29,...,40", "50,51,...,85", 5e-3),
    ("+proj=ob_tran +o_proj=longlat +lon_0=0 +o_lat_p=25 +R=6.371e+06 +no_defs", "-46.4,-46.2,...,46.", "-36.4,-36.2,...,38.8",
     "degree", "-90,-89,...,90", "40,51,...,85", 3e-2),
    ("+proj=latlon +R=6.371e+06 +no_defs", "-179,-178,...,179", "-89.5,-89,...,89.5", "degree", "-180,-179,...,179",
     "-90,-89,...,90", 1e-3),
]


@pytest.mark.parametrize("case", range(len(_BACKFORTH)))
@pytest.mark.parametrize("wind", [(0.0, 1.0), (1.0, 0.0)], ids=["north", "east"])
def test_interpolator_vector_backforth(case, wind):
    proj, x_axis, y_axis, unit, lon_axis, lat_axis, delta = _BACKFORTH[case]
    lon = np.arange(-180.0, 180.0, 1.0)
    lat = np.arange(-90.0, 90.5, 1.0)
    xw = np.full((lat.size, lon.size), wind[0], np.float32)
    yw = np.full((lat.size, lon.size), wind[1], np.float32)

    interp = fb.Interpolator("+proj=latlong +R=" + _R + " +no_defs", lon, lat, True, has_xy_vectors=True)
    interp.changeProjection(fb.Method.NEAREST_NEIGHBOR, proj, x_axis, y_axis, unit, unit)
    x1 = interp.getDataSlice(xw, counterpart=yw, direction="x")
    y1 = interp.getDataSlice(yw, counterpart=xw, direction="y")
    assert np.isfinite(x1).any()

    is_degree = unit == "degree"
    iback = fb.Interpolator(proj, fb.spatial_axis_spec(x_axis), fb.spatial_axis_spec(y_axis), is_degree, has_xy_vectors=True)
    iback.changeProjection(fb.Method.NEAREST_NEIGHBOR, "+proj=latlon +R=" + _R, lon_axis, lat_axis, "degrees_east", "degrees_north")
    x2 = iback.getDataSlice(x1, counterpart=y1, direction="x")
    y2 = iback.getDataSlice(y1, counterpart=x1, direction="y")

    # getScaledData (CDMReader.cc scaleDataOf) turns the variable's fill value back into NaN before the comparison
    fill = fb.default_fill_value(np.float32)
    ok = ~((x2 == np.float32(fill)) | (y2 == np.float32(fill)) | np.isnan(x2) | np.isnan(y2))
    assert ok.sum() > 0.1 * ok.size, (int(ok.sum()), ok.size)  # hirlam8 covers only 338 of the 2556 lat/lon points
    assert np.abs(x2[ok] - wind[0]).max() < delta, float(np.abs(x2[ok] - wind[0]).max())
    assert np.abs(y2[ok] - wind[1]).max() < delta, float(np.abs(y2[ok] - wind[1]).max())


def test_vector_pair_cache_serves_the_counterpart_without_a_second_pass():
    # the reference interpolates and rotates both components on each component's call
    # (CDMInterpolator.cc:259-276); with a pair key the second request is served from the parked half
    lon, lat, _ = _config2_like(40)
    rng = np.random.default_rng(21)
    u = rng.normal(0, 10, (3, lat.size, lon.size)).astype(np.float32)
    v = rng.normal(0, 10, (3, lat.size, lon.size)).astype(np.float32)
    ax = (np.arange(50) - 24.5) * 0.5
    interp = fb.Interpolator(SRC_LL, lon, lat, True, has_xy_vectors=True)
    interp.changeProjection("bilinear", ROTPOLE, ax, ax, "degree", "degree")
    plain_x = interp.getDataSlice(u, counterpart=v, direction="x")
    plain_y = interp.getDataSlice(v, counterpart=u, direction="y")
    n0 = fb.kernel_launches()
    got_x = interp.getDataSlice(u, counterpart=v, direction="x", pair_key=("x_wind", "y_wind", 0))
    n1 = fb.kernel_launches()
    got_y = interp.getDataSlice(v, counterpart=u, direction="y", pair_key=("x_wind", "y_wind", 0))
    n2 = fb.kernel_launches()
    assert n1 > n0 and n2 == n1  # the second half launched nothing
    assert_bit_equal(got_x, plain_x, "pair cache x")
    assert_bit_equal(got_y, plain_y, "pair cache y")
    # a parked half is handed out once; another slice (key) is computed afresh; y first works the same way
    got_y2 = interp.getDataSlice(v, counterpart=u, direction="y", pair_key=("x_wind", "y_wind", 1))
    assert fb.kernel_launches() > n2
    n3 = fb.kernel_launches()
    got_x2 = interp.getDataSlice(u, counterpart=v, direction="x", pair_key=("x_wind", "y_wind", 1))
    assert fb.kernel_launches() == n3
    assert_bit_equal(got_x2, plain_x, "pair cache x after y")
    assert_bit_equal(got_y2, plain_y, "pair cache y first")
    for k in range(10):  # bounded
        interp.getDataSlice(u, counterpart=v, direction="x", pair_key=("x_wind", "y_wind", 100 + k))
    assert len(interp._pair_cache) <= interp.pairCacheSlots
    interp.changeProjection("bilinear", ROTPOLE, ax, ax, "degree", "degree")
    assert not interp._pair_cache


@pytest.mark.parametrize("out_dtype", [np.int16, np.int32, np.uint8, np.uint16])
def test_typed_store_rounding_is_exhaustive_per_exponent(oracle, out_dtype):
    # interpolationArray2Data for integer targets is lround() narrowed to int (include/fimex/Utils.h:88-113, 444-464).  The
    # staged store rounds values below 2^21 on the integer pipe and larger ones in fp64: every mantissa of a set of exponents
    # either side of that switch, both signs, through the nearest-neighbour gather with an identity table
    nx, ny = 4096, 2048  # 2^23 points per level: one exponent per level
    px = np.tile(np.arange(nx, dtype=np.float64), ny)
    py = np.repeat(np.arange(ny, dtype=np.float64), nx)
    ci = fb.CachedInterpolation("x", "y", Method.NEAREST_NEIGHBOR, px, py, nx, ny, nx, ny)
    mant = np.arange(1 << 23, dtype=np.uint32)
    fill = fb.default_fill_value(out_dtype)
    for exps in ((-3, -2, -1, 0), (1, 2, 7, 14), (15, 16, 19, 20), (21, 22, 23, 30)):
        for sign in (0, 0x80000000):
            bits = np.stack([((np.uint32(e + 127) << np.uint32(23)) | mant | np.uint32(sign)) for e in exps])
            v = bits.view(np.float32).reshape(len(exps), ny, nx)
            got = ci.getDataSlice(v, fill, outType=out_dtype)
            want = oracle.from_float(v, fill, out_dtype)
            assert np.array_equal(got.ravel(), want.ravel()), (exps, sign, out_dtype)
    # NaN -> fill, zeros, ties and the float neighbours of the ties
    ties = np.array([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 0.49999997, 0.50000006, -0.49999997, 1048575.5, 2097151.5, 2097152.0, 2097152.5, -2097151.5,
                     4194303.5, 8388607.5, 0.0, -0.0, np.nan, 32767.5, -32768.5, 65535.5, 1e-45, -1e-45], np.float32)
    v = np.zeros((1, ny, nx), np.float32)
    v.ravel()[:ties.size] = ties
    got = ci.getDataSlice(v, fill, outType=out_dtype)
    assert np.array_equal(got.ravel(), oracle.from_float(v, fill, out_dtype).ravel())


def test_bicubic_contracted_mode_is_opt_in_and_within_tolerance(oracle, monkeypatch):
    # FIMEX_B200_BICUBIC_CONTRACT=1 evaluates the bicubic sums as fp64 FMA chains with one final rounding to float (20 fp64
    # instructions per value instead of 35): not bit-identical, but inside the 1e-5 relative bar the north star states for
    # interpolated floats.  Default (unset) stays bit-identical to the reference.
    lon, lat, ax = _config2_like(300)
    rng = np.random.default_rng(8)
    field = rng.normal(250, 30, (9, lat.size, lon.size)).astype(np.float32)
    u = rng.normal(0, 12, (3, lat.size, lon.size)).astype(np.float32)
    v = rng.normal(0, 12, (3, lat.size, lon.size)).astype(np.float32)
    ci = fb.CachedInterpolation.fromProjection(Method.BICUBIC, ROTPOLE, ax, ax, True, True, SRC_LL, lon, lat, True)
    cvr = fb.CachedVectorReprojection.fromProjection(fb.MIFI_VECTOR_KEEP_SIZE, SRC_LL, ROTPOLE, ax, ax, fb.LONGITUDE, fb.LATITUDE)
    monkeypatch.delenv("FIMEX_B200_BICUBIC_CONTRACT", raising=False)
    exact = ci.interpolateValues(field)
    eu, ev = ci.interpolateVector(u, v, cvr)
    gx, gy = ci.points()
    assert_bit_equal(exact, oracle.cached_interpolate(2, gx, gy, lon.size, lat.size, ax.size, ax.size, field), "bicubic exact (default)")
    monkeypatch.setenv("FIMEX_B200_BICUBIC_CONTRACT", "1")
    fast = ci.interpolateValues(field)
    fu, fv = ci.interpolateVector(u, v, cvr)
    monkeypatch.delenv("FIMEX_B200_BICUBIC_CONTRACT")
    assert np.array_equal(np.isnan(fast), np.isnan(exact))
    ok = ~np.isnan(exact)
    assert ok.sum() > 0.5 * ok.size
    assert np.abs(fast[ok] - exact[ok]).max() <= 1e-5 * np.abs(field).max()
    assert (fast[ok] != exact[ok]).any()  # the mode was really taken
    for f, e in ((fu, eu), (fv, ev)):
        m = ~np.isnan(e)
        assert np.array_equal(np.isnan(f), ~m) and np.abs(f[m] - e[m]).max() <= 1e-5 * 60.0
    assert_bit_equal(ci.interpolateValues(field), exact, "bicubic exact again")


def test_interpolator_device_resident_slices():
    # the same getDataSlice with CUDA tensors: crop, adapters, gather and rotation on the device, CUDA tensors back
    import torch
    lon, lat, _ = _config2_like(40)
    rng = np.random.default_rng(31)
    t = rng.integers(-3000, 3000, (4, lat.size, lon.size)).astype(np.int16)
    t[rng.random(t.shape) < 0.03] = -32767
    u = rng.normal(0, 10, (2, lat.size, lon.size)).astype(np.float32)
    v = rng.normal(0, 10, (2, lat.size, lon.size)).astype(np.float32)
    ax = (np.arange(64) - 31.5) * 0.4
    interp = fb.Interpolator(SRC_LL, lon, lat, True, has_xy_vectors=True)
    interp.changeProjection("bilinear", ROTPOLE, ax, ax, "degree", "degree")
    want_t = interp.getDataSlice(t)
    want_u = interp.getDataSlice(u, counterpart=v, direction="x")
    want_v = interp.getDataSlice(v, counterpart=u, direction="y")
    dt, du, dv = (torch.from_numpy(a).cuda() for a in (t, u, v))
    got_t = interp.getDataSlice(dt)
    assert got_t.is_cuda and got_t.dtype == torch.int16 and tuple(got_t.shape) == want_t.shape
    assert np.array_equal(got_t.cpu().numpy(), want_t)
    got_u = interp.getDataSlice(du, counterpart=dv, direction="x", pair_key=("u", "v", 0))
    got_v = interp.getDataSlice(dv, counterpart=du, direction="y", pair_key=("u", "v", 0))
    assert got_u.is_cuda and got_v.is_cuda
    assert_bit_equal(got_u.cpu().numpy(), want_u, "device-resident x component")
    assert_bit_equal(got_v.cpu().numpy(), want_v, "device-resident y component")


def test_concurrent_large_pageable_slices(oracle):
    """Calls big enough to be cut into chunks and staged through page-locked bounce buffers (pageable numpy arrays in and out,
    > 64 MB per call), from four host threads on one handle: the chunk pipeline, the pinned cache and the copy threads are
    all per call or locked."""
    import threading
    lon, lat, ax = _config2_like(700)
    ci = fb.CachedInterpolation.fromProjection(Method.BILINEAR, ROTPOLE, ax, ax, True, True, SRC_LL, lon, lat, True)
    ci.createReducedDomain()
    rng = np.random.default_rng(5)
    nz = 60  # 60 x 700 x 700 x 4 B = 118 MB out per call
    fields = [rng.normal(250, 30, (nz, ci.getInY(), ci.getInX())).astype(np.float32) for _ in range(4)]
    gx, gy = ci.points()
    want = [oracle.cached_interpolate(1, gx, gy, ci.getInX(), ci.getInY(), ax.size, ax.size, f) for f in fields]
    got = [None] * 4
    errors = []

    def work(i):
        try:
            for _ in range(2):
                got[i] = ci.interpolateValues(fields[i])
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i in range(4):
        assert_bit_equal(got[i], want[i], f"thread {i}")


@pytest.mark.parametrize("mode,levels", [("1", "8"), ("1", "4"), ("2", "8"), ("2", "4")])
def test_bulk_store_forms_are_bit_identical(oracle, monkeypatch, mode, levels):
    """FIMEX_B200_BULK_STORE=1 (cp.async.bulk row copies) / 2 (cp.async.bulk.tensor boxes): the bilinear gather with its output tile
    staged in shared memory and stored by the copy engine -- opt-in (slower than per-thread stores on B200,
    profiles/r02_bulk_store_ab.txt), same bits: float, fill-value float and int16 output, partial tiles, partial batches, a
    many-tap tile (target coarser than the source) handled by the fallback launch"""
    monkeypatch.setenv("FIMEX_B200_BULK_STORE", mode)
    monkeypatch.setenv("FIMEX_B200_BULK_LEVELS", levels)
    for (inX, inY, inZ, outX, outY, angle, zoom) in ((60, 50, 19, 200, 152, 17.0, 5.0), (90, 70, 70, 328, 77, -33.0, 6.5), (300, 40, 9, 256, 96, 3.0, 0.6)):
        px, py = _smooth_positions(inX, inY, outX, outY, angle, zoom, 7)
        rng = np.random.default_rng(inZ)
        field = rng.normal(250, 30, (inZ, inY, inX)).astype(np.float32)
        field[rng.random(field.shape) < 0.02] = np.nan
        want = oracle.cached_interpolate(1, px, py, inX, inY, outX, outY, field)
        ci = fb.CachedInterpolation("x", "y", Method.BILINEAR, px, py, inX, inY, outX, outY)
        assert_bit_equal(ci.interpolateValues(field), want, f"bulk store mode {mode}, {levels} levels, plain float")
        fill = np.float32(9.96921e+36)
        filled = np.where(np.isnan(field), fill, field)
        got = ci.getDataSlice(filled, float(fill))
        assert_bit_equal(got, np.where(np.isnan(want), fill, want), f"bulk store mode {mode}: float with fill values")
        packed = np.clip(np.round(np.nan_to_num(field, nan=250.0) * 10), -32000, 32000).astype(np.int16)
        packed[np.isnan(field)] = -32767
        want16 = oracle.from_float(oracle.cached_interpolate(1, px, py, inX, inY, outX, outY, oracle.as_float(packed, -32767.0)), -32767.0, np.int16)
        got16 = ci.getDataSlice(packed, -32767.0)
        assert got16.dtype == np.int16 and np.array_equal(got16, want16), f"bulk store mode {mode}: int16"


def _stencil_absmax(field, gx, gy):
    """max |tap| over the 4 x 4 bicubic stencil of every target point (NaN where the stencil leaves the grid): the scale the
    fp32 mode's error is measured against"""
    nz, iy, ix = field.shape
    a = np.abs(np.nan_to_num(field, nan=0.0)).max(axis=0)
    win = np.lib.stride_tricks.sliding_window_view(a, (4, 4)).max(axis=(2, 3))  # [iy-3][ix-3], window starting at (y, x)
    x0, y0 = np.floor(gx).astype(np.int64) - 1, np.floor(gy).astype(np.int64) - 1
    ok = (x0 >= 0) & (x0 + 3 < ix) & (y0 >= 0) & (y0 + 3 < iy)
    d = np.full(gx.shape, np.nan)
    d[ok] = win[y0[ok], x0[ok]]
    return d


def test_bicubic_fp32_mode_is_opt_in_and_within_tolerance(oracle, monkeypatch):
    """FIMEX_B200_BICUBIC_FP32=1: weights computed in fp64 as the reference does, rounded to fp32 once, 20 fp32 FMAs per output.
    Default stays bit-identical; the mode keeps the NaN mask exactly and every value within 1e-5 of the largest |tap| of the
    point's own 4 x 4 stencil (north_star: <= 1e-5 relative for bicubic) -- scalar, fill-value and int16 slices, u/v with rotation."""
    lon, lat, ax = _config2_like(300)
    rng = np.random.default_rng(18)
    field = rng.normal(250, 30, (19, lat.size, lon.size)).astype(np.float32)
    field[rng.random(field.shape) < 0.002] = np.nan
    u = rng.normal(0, 12, (5, lat.size, lon.size)).astype(np.float32)
    v = rng.normal(0, 12, (5, lat.size, lon.size)).astype(np.float32)
    ci = fb.CachedInterpolation.fromProjection(Method.BICUBIC, ROTPOLE, ax, ax, True, True, SRC_LL, lon, lat, True)
    cvr = fb.CachedVectorReprojection.fromProjection(fb.MIFI_VECTOR_KEEP_SIZE, SRC_LL, ROTPOLE, ax, ax, fb.LONGITUDE, fb.LATITUDE)
    monkeypatch.delenv("FIMEX_B200_BICUBIC_FP32", raising=False)
    monkeypatch.delenv("FIMEX_B200_BICUBIC_CONTRACT", raising=False)
    gx, gy = ci.points()
    exact = ci.interpolateValues(field)
    assert_bit_equal(exact, oracle.cached_interpolate(2, gx, gy, lon.size, lat.size, ax.size, ax.size, field), "bicubic exact (default)")
    eu, ev = ci.interpolateVector(u, v, cvr)
    monkeypatch.setenv("FIMEX_B200_BICUBIC_FP32", "1")
    fast = ci.interpolateValues(field)
    fu, fv = ci.interpolateVector(u, v, cvr)
    fill = np.float32(9.96921e+36)
    fast_fill = ci.getDataSlice(np.where(np.isnan(field), fill, field), float(fill))
    monkeypatch.delenv("FIMEX_B200_BICUBIC_FP32")
    assert np.array_equal(np.isnan(fast), np.isnan(exact))
    ok = ~np.isnan(exact)
    assert ok.sum() > 0.5 * ok.size
    scale = _stencil_absmax(field, gx, gy).reshape(1, ax.size, ax.size)
    err = np.abs(fast.astype(np.float64) - exact.astype(np.float64)) / scale
    assert np.nanmax(err[ok]) <= 1e-5, np.nanmax(err[ok])
    assert (fast[ok] != exact[ok]).any()  # the mode was really taken
    assert np.array_equal(fast_fill == fill, ~ok) and np.array_equal(fast_fill[ok], fast[ok])  # the adapters around it are unchanged
    su = np.maximum(_stencil_absmax(u, gx, gy), _stencil_absmax(v, gx, gy)).reshape(1, ax.size, ax.size)
    for f, e in ((fu, eu), (fv, ev)):
        m = ~np.isnan(e)
        assert np.array_equal(np.isnan(f), ~m)
        assert np.nanmax((np.abs(f.astype(np.float64) - e.astype(np.float64)) / su)[m]) <= 2e-5  # two components enter each rotated value
    assert_bit_equal(ci.interpolateValues(field), exact, "bicubic exact again")


def test_bilinear_quad_layout_is_bit_identical(oracle, monkeypatch):
    """FIMEX_B200_BILINEAR_QUAD=1 (read when the tables are built): a thread owns 4 x-neighbours and re-uses the taps of the
    previous point -- opt-in (slower on B200, profiles/r02_bilinear_quad_ab.txt), same bits: scalar, fill values, int16, u/v with
    rotation, partial tiles and batches, a many-tap tile"""
    monkeypatch.setenv("FIMEX_B200_BILINEAR_QUAD", "1")
    for (inX, inY, inZ, outX, outY, angle, zoom) in ((60, 50, 19, 200, 152, 17.0, 5.0), (90, 70, 70, 331, 77, -33.0, 6.5), (300, 40, 9, 256, 96, 3.0, 0.6)):
        px, py = _smooth_positions(inX, inY, outX, outY, angle, zoom, 11)
        rng = np.random.default_rng(inZ + 1)
        field = rng.normal(250, 30, (inZ, inY, inX)).astype(np.float32)
        field[rng.random(field.shape) < 0.02] = np.nan
        want = oracle.cached_interpolate(1, px, py, inX, inY, outX, outY, field)
        ci = fb.CachedInterpolation("x", "y", Method.BILINEAR, px, py, inX, inY, outX, outY)
        assert_bit_equal(ci.interpolateValues(field), want, "quad layout, plain float")
        fill = np.float32(9.96921e+36)
        got = ci.getDataSlice(np.where(np.isnan(field), fill, field), float(fill))
        assert_bit_equal(got, np.where(np.isnan(want), fill, want), "quad layout: float with fill values")
        packed = np.clip(np.round(np.nan_to_num(field, nan=250.0) * 10), -32000, 32000).astype(np.int16)
        packed[np.isnan(field)] = -32767
        want16 = oracle.from_float(oracle.cached_interpolate(1, px, py, inX, inY, outX, outY, oracle.as_float(packed, -32767.0)), -32767.0, np.int16)
        assert np.array_equal(ci.getDataSlice(packed, -32767.0), want16), "quad layout: int16"
        m = np.zeros((outX * outY, 4))
        ang = rng.uniform(0, 2 * np.pi, outX * outY)
        m[:, 0], m[:, 1], m[:, 2], m[:, 3] = np.cos(ang), np.sin(ang), -np.sin(ang), ang
        cvr = fb.CachedVectorReprojection(fb.MIFI_VECTOR_KEEP_SIZE, m.ravel(), outX, outY)
        v = rng.normal(0, 10, field.shape).astype(np.float32)
        gu, gv = ci.interpolateVector(np.nan_to_num(field, nan=1.0), v, cvr)
        wu, wv = oracle.vector_reproject_by_matrix(m.ravel(), oracle.cached_interpolate(1, px, py, inX, inY, outX, outY, np.nan_to_num(field, nan=1.0)),
                                                   oracle.cached_interpolate(1, px, py, inX, inY, outX, outY, v), outX, outY, inZ)
        assert_bit_equal(gu, wu, "quad layout: rotated u")
        assert_bit_equal(gv, wv, "quad layout: rotated v")


@pytest.mark.parametrize("tma", ["0", "1"])
def test_bicubic_store_paths_are_bit_identical(oracle, monkeypatch, tma):
    """the bicubic gather's two ways out of its shared-memory output tile -- copy engine (cp.async.bulk.tensor.3d, default for
    float output whose rows are 16-byte aligned) and per-thread stores (FIMEX_B200_BICUBIC_TMA=0; always for other types and
    unaligned rows): same bits, scalar / fill values / u, v with rotation, whole and partial tiles, partial batches"""
    monkeypatch.setenv("FIMEX_B200_BICUBIC_TMA", tma)
    monkeypatch.delenv("FIMEX_B200_BICUBIC_FP32", raising=False)
    for (inX, inY, inZ, outX, outY, angle, zoom) in ((60, 50, 19, 200, 152, 17.0, 5.0), (90, 70, 70, 332, 77, -33.0, 6.5), (64, 48, 9, 100, 61, 40.0, 3.0)):
        px, py = _smooth_positions(inX, inY, outX, outY, angle, zoom, 5)
        rng = np.random.default_rng(inZ + 7)
        field = rng.normal(250, 30, (inZ, inY, inX)).astype(np.float32)
        field[rng.random(field.shape) < 0.01] = np.nan
        want = oracle.cached_interpolate(2, px, py, inX, inY, outX, outY, field)
        ci = fb.CachedInterpolation("x", "y", Method.BICUBIC, px, py, inX, inY, outX, outY)
        assert_bit_equal(ci.interpolateValues(field), want, f"bicubic TMA={tma}, plain float")
        fill = np.float32(9.96921e+36)
        got = ci.getDataSlice(np.where(np.isnan(field), fill, field), float(fill))
        assert_bit_equal(got, np.where(np.isnan(want), fill, want), f"bicubic TMA={tma}: float with fill values")
        m = np.zeros((outX * outY, 4))
        ang = rng.uniform(0, 2 * np.pi, outX * outY)
        m[:, 0], m[:, 1], m[:, 2], m[:, 3] = np.cos(ang), np.sin(ang), -np.sin(ang), ang
        cvr = fb.CachedVectorReprojection(fb.MIFI_VECTOR_KEEP_SIZE, m.ravel(), outX, outY)
        u = np.nan_to_num(field, nan=1.0)
        v = rng.normal(0, 10, field.shape).astype(np.float32)
        gu, gv = ci.interpolateVector(u, v, cvr)
        wu, wv = oracle.vector_reproject_by_matrix(m.ravel(), oracle.cached_interpolate(2, px, py, inX, inY, outX, outY, u),
                                                   oracle.cached_interpolate(2, px, py, inX, inY, outX, outY, v), outX, outY, inZ)
        assert_bit_equal(gu, wu, f"bicubic TMA={tma}: rotated u")
        assert_bit_equal(gv, wv, f"bicubic TMA={tma}: rotated v")
        pu, pv = ci.interpolateVector(u, v, None)
        assert_bit_equal(pu, oracle.cached_interpolate(2, px, py, inX, inY, outX, outY, u), f"bicubic TMA={tma}: u, no rotation")


@pytest.mark.parametrize("mode", ["1", "2"])
def test_nearest_neighbour_bulk_store_forms_are_bit_identical(oracle, monkeypatch, mode):
    """FIMEX_B200_NN_BULK=1 (cp.async.bulk row copies) / 2 (cp.async.bulk.tensor boxes): the nearest-neighbour gather with its output
    tile in shared memory -- opt-in (slower on B200, profiles/r02_nn_bulk_ab.txt), same bits, NaN payload included: plain,
    fill values, int32 output, partial tiles and batches, a many-tap tile"""
    monkeypatch.setenv("FIMEX_B200_NN_BULK", mode)
    for (inX, inY, inZ, outX, outY, angle, zoom) in ((60, 50, 19, 200, 152, 17.0, 5.0), (90, 70, 70, 332, 77, -33.0, 6.5), (300, 40, 9, 256, 96, 3.0, 0.6)):
        px, py = _smooth_positions(inX, inY, outX, outY, angle, zoom, 3)
        rng = np.random.default_rng(inZ + 2)
        field = rng.normal(250, 30, (inZ, inY, inX)).astype(np.float32)
        field[rng.random(field.shape) < 0.02] = np.nan
        want = oracle.cached_interpolate(0, px, py, inX, inY, outX, outY, field)
        ci = fb.CachedInterpolation("x", "y", Method.NEAREST_NEIGHBOR, px, py, inX, inY, outX, outY)
        assert_bit_equal(ci.interpolateValues(field), want, f"NN bulk mode {mode}", nan_payload=True)
        fill = np.float32(9.96921e+36)
        got = ci.getDataSlice(np.where(np.isnan(field), fill, field), float(fill))
        assert_bit_equal(got, np.where(np.isnan(want), fill, want), f"NN bulk mode {mode}: fill values")
        ints = rng.integers(-100000, 100000, field.shape).astype(np.int32)
        ints[rng.random(field.shape) < 0.02] = -2147483647
        want32 = oracle.from_float(oracle.cached_interpolate(0, px, py, inX, inY, outX, outY, oracle.as_float(ints, -2147483647.0)), -2147483647.0, np.int32)
        assert np.array_equal(ci.getDataSlice(ints, -2147483647.0), want32), f"NN bulk mode {mode}: int32"
