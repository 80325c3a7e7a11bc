"""A10 / P: the fp64 coordinate transforms against INDEPENDENT known answers (tests/proj_known_answers.py): Snyder's printed
numerical examples, Snyder's equations written out in numpy, and the rotated pole as a 3-D change of basis.  Each case is
asserted twice: against the CPU oracle (oracle/pj_oracle.c, `-m "not gpu"`) and against the GPU through the C ABI
`mifi_project_values` / `mifi_project_axes` (`-m gpu`) -- so the two restatements are pinned to published numbers, not to
each other.  Replaces: src/interpolation.c:1158-1244 -> pj_transform (PROJ.4, absent from /root/reference).
Bars: the printed digits for the printed examples; <= 1e-9 degree (angles) / <= 1e-9 degree of arc x a (metres) elsewhere.
"""
import numpy as np
import pytest

import proj_known_answers as ka
from proj_known_answers import DEG

ARC = 1e-9 * DEG  # 1e-9 degree in radians


class _OracleTransform:
    name = "oracle"

    def __init__(self, oracle):
        self.o = oracle

    def values(self, pin, pout, x, y):
        rc, a, b = self.o.project_values(pin, pout, x, y)
        assert rc == 1
        return a, b

    def axes(self, pin, pout, xa, ya):
        rc, a, b = self.o.project_axes(pin, pout, xa, ya)
        assert rc == 1
        return a, b


class _GpuTransform:
    name = "gpu"

    def __init__(self):
        import fimex_b200 as fb
        self.fb = fb

    def values(self, pin, pout, x, y):
        rc, a, b = self.fb.mifi_project_values(pin, pout, np.array(x, dtype=np.float64), np.array(y, dtype=np.float64))
        assert rc == self.fb.MIFI_OK
        return a, b

    def axes(self, pin, pout, xa, ya):
        rc, a, b = self.fb.mifi_project_axes(pin, pout, np.array(xa, dtype=np.float64), np.array(ya, dtype=np.float64))
        assert rc == self.fb.MIFI_OK
        return a, b


@pytest.fixture(params=["oracle", pytest.param("gpu", marks=pytest.mark.gpu)])
def tr(request):
    if request.param == "oracle":
        return _OracleTransform(request.getfixturevalue("oracle"))
    return _GpuTransform()


@pytest.mark.parametrize("case", ka.SNYDER, ids=[c[0] for c in ka.SNYDER])
def test_snyder_printed_examples(tr, case):
    name, proj, lon, lat, x, y, tol = case
    ll = ka.LATLONG_FOR[name]
    gx, gy = tr.values(ll, proj, [lon * DEG], [lat * DEG])
    assert abs(gx[0] - x) <= tol and abs(gy[0] - y) <= tol, (tr.name, name, gx[0], gy[0])
    # inverse of the printed (rounded) x, y: back to the printed lon/lat within what the rounding of x, y allows
    blon, blat = tr.values(proj, ll, [x], [y])
    scale = 1.0 if "R=1" in proj else 6378388.0
    back_tol = 4 * tol / scale + ARC
    assert ka.angle_diff(blon[0], lon * DEG) <= back_tol and abs(blat[0] - lat * DEG) <= back_tol, (tr.name, name, blon[0] / DEG, blat[0] / DEG)


def _random_lonlat(rng, n, lat_lo, lat_hi):
    return rng.uniform(-180, 180, n) * DEG, rng.uniform(lat_lo, lat_hi, n) * DEG


@pytest.mark.parametrize("figure", ["sphere", "WGS84", "clrk66"])
@pytest.mark.parametrize("pars", [(33.0, 45.0, 23.0, -96.0), (63.0, 63.0, 63.0, 15.0), (-30.0, -60.0, -45.0, 140.0), (77.5, 77.5, 77.5, -25.0)])
def test_lcc_against_snyder_equations(tr, figure, pars):
    """forward on 2e4 random points against eqs. 15-7..15-10 / 14-1, 14-2, then the transform's own inverse back"""
    lat1, lat2, lat0, lon0 = pars
    a, es, fig = {"sphere": (6371000.0, 0.0, "+R=6371000"), "WGS84": (6378137.0, 2 / 298.257223563 - 298.257223563**-2, "+ellps=WGS84"),
                  "clrk66": (6378206.4, 1 - (6356583.8 / 6378206.4)**2, "+ellps=clrk66")}[figure]
    proj = f"+proj=lcc +lat_1={lat1} +lat_2={lat2} +lat_0={lat0} +lon_0={lon0} {fig} +no_defs"
    ll = f"+proj=latlong {fig} +no_defs"
    rng = np.random.default_rng(15)
    south = lat1 < 0
    lon, lat = _random_lonlat(rng, 20000, -85 if south else -60, 60 if south else 85)
    lon = lon0 * DEG + (lon - lon0 * DEG + np.pi) % (2 * np.pi) - np.pi  # stay on the cone's own 360 degrees, away from the cut
    lon = lon0 * DEG + 0.98 * (lon - lon0 * DEG)
    wx, wy = ka.lcc_forward(lon, lat, lat1 * DEG, lat2 * DEG, lat0 * DEG, lon0 * DEG, a, es)
    gx, gy = tr.values(ll, proj, lon, lat)
    rho = np.hypot(wx, wy) + a  # a point far down the cone is many earth radii out: the bar is 1e-9 degree of arc at its radius
    assert (np.abs(gx - wx) <= ARC * rho).all() and (np.abs(gy - wy) <= ARC * rho).all(), (tr.name, np.abs(gx - wx).max(), np.abs(gy - wy).max())
    blon, blat = tr.values(proj, ll, wx, wy)
    assert ka.angle_diff(blon, lon).max() <= ARC and np.abs(blat - lat).max() <= ARC, (tr.name, ka.angle_diff(blon, lon).max() / DEG)


@pytest.mark.parametrize("figure", ["sphere", "WGS84", "intl"])
@pytest.mark.parametrize("pars", [(90.0, 60.0, 0.0), (90.0, 90.0, -32.0), (-90.0, -71.0, -100.0), (90.0, 70.0, 58.0)])
def test_polar_stereographic_against_snyder_equations(tr, figure, pars):
    """eqs. 21-30, 21-31, 21-34 (ellipsoid; sphere for e = 0) on random points of the pole's hemisphere, and the inverse"""
    lat0, lat_ts, lon0 = pars
    a, es, fig = {"sphere": (6371000.0, 0.0, "+a=6371000 +e=0"), "WGS84": (6378137.0, 2 / 298.257223563 - 298.257223563**-2, "+ellps=WGS84"),
                  "intl": (6378388.0, 2 / 297.0 - 297.0**-2, "+ellps=intl")}[figure]
    proj = f"+proj=stere +lat_0={lat0} +lat_ts={lat_ts} +lon_0={lon0} {fig} +no_defs"
    ll = f"+proj=latlong {fig} +no_defs"
    rng = np.random.default_rng(21)
    south = lat0 < 0
    lon, lat = _random_lonlat(rng, 20000, -89.9 if south else 10, -10 if south else 89.9)
    if abs(lat_ts) == 90.0:  # eq. 21-33 (true scale at the pole): rho = 2 a k0 t / [(1+e)^(1+e) (1-e)^(1-e)]^(1/2), k0 = 1
        e = np.sqrt(es)
        sgn = -1.0 if south else 1.0
        rho = 2 * a * ka._t(sgn * lat, e) / np.sqrt((1 + e)**(1 + e) * (1 - e)**(1 - e))
        dl = sgn * (lon - lon0 * DEG)
        wx, wy = sgn * rho * np.sin(dl), -sgn * rho * np.cos(dl)
    else:
        wx, wy = ka.stere_polar_forward(lon, lat, lat_ts * DEG, lon0 * DEG, a, es, south)
    gx, gy = tr.values(ll, proj, lon, lat)
    rho = np.hypot(wx, wy) + a
    assert (np.abs(gx - wx) <= ARC * rho).all() and (np.abs(gy - wy) <= ARC * rho).all(), (tr.name, np.abs(gx - wx).max(), np.abs(gy - wy).max())
    blon, blat = tr.values(proj, ll, wx, wy)
    assert ka.angle_diff(blon, lon).max() <= ARC and np.abs(blat - lat).max() <= ARC


@pytest.mark.parametrize("pars", [(40.0, -100.0, 1.0), (0.0, 10.0, 0.9996), (-35.0, 150.0, 1.0), (63.0, 15.0, 1.0)])
def test_oblique_stereographic_sphere_against_snyder_equations(tr, pars):
    lat0, lon0, k0 = pars
    R = 6371000.0
    proj = f"+proj=stere +lat_0={lat0} +lon_0={lon0} +k={k0} +R={R} +no_defs"
    ll = f"+proj=latlong +R={R} +no_defs"
    rng = np.random.default_rng(4)
    # points within 120 degrees of the centre (the antipode is the projection's singularity)
    c = ka._unit(lon0 * DEG, lat0 * DEG)
    lon, lat = _random_lonlat(rng, 40000, -89.9, 89.9)
    keep = ka._unit(lon, lat) @ c > np.cos(120 * DEG)
    lon, lat = lon[keep], lat[keep]
    wx, wy = ka.stere_oblique_sphere_forward(lon, lat, lat0 * DEG, lon0 * DEG, R, k0)
    gx, gy = tr.values(ll, proj, lon, lat)
    rho = np.hypot(wx, wy) + R
    assert (np.abs(gx - wx) <= ARC * rho).all() and (np.abs(gy - wy) <= ARC * rho).all()
    blon, blat = tr.values(proj, ll, wx, wy)
    assert ka.angle_diff(blon, lon).max() <= ARC and np.abs(blat - lat).max() <= ARC


@pytest.mark.parametrize("pars", [(22.0, -40.0), (25.0, 0.0), (37.5, 177.5), (-10.0, 12.0), (89.0, -40.0)])
def test_rotated_pole_against_change_of_basis(tr, pars):
    """ob_tran +o_proj=longlat (the target projection of BASELINE config 2: +lon_0=-40 +o_lat_p=22) on 1e5 random points of
    the whole sphere against an explicit 3-D rotation, both directions: longitude and latitude <= 1e-12 degree away from the
    poles of both frames (80 degrees), position on the sphere <= 2e-11 degree of arc everywhere (asin and atan2 are
    ill-conditioned at the poles: one ulp of their argument is amplified 200 x at 89.9 degrees)"""
    o_lat_p, lon_0 = pars
    proj = f"+proj=ob_tran +o_proj=longlat +lon_0={lon_0} +o_lat_p={o_lat_p} +R=6.371e+06 +no_defs"
    ll = "+proj=latlong +R=6.371e+06 +no_defs"
    rng = np.random.default_rng(22)
    n = 100_000
    lon = rng.uniform(-np.pi, np.pi, n)
    lat = np.arcsin(rng.uniform(-1, 1, n))
    bar = 1e-12 * DEG
    # geographic -> rotated
    wl, wp = ka.geographic_to_rotated(lon, lat, o_lat_p * DEG, lon_0 * DEG)
    gl, gp = tr.values(ll, proj, lon, lat)
    # asin near +-90 degrees loses half the digits in ANY implementation: compare positions on the sphere (arc length)
    arc = np.linalg.norm(ka._unit(gl, gp) - ka._unit(wl, wp), axis=-1)
    assert arc.max() <= 20 * bar, (tr.name, "fwd", arc.max() / DEG)  # asin(1 - 1e-5) amplifies one ulp of its argument 200 x
    away = np.abs(wp) < 89.9 * DEG
    assert ka.angle_diff(gl[away], wl[away]).max() <= 600 * bar and np.abs(gp[away] - wp[away]).max() <= 600 * bar  # 1/cos(89.9 deg) = 573
    mid = (np.abs(wp) < 80 * DEG) & (np.abs(lat) < 80 * DEG)  # away from the poles of BOTH frames: the 1e-12 degree bar itself
    assert ka.angle_diff(gl[mid], wl[mid]).max() <= bar and np.abs(gp[mid] - wp[mid]).max() <= bar
    # rotated -> geographic (the direction the index tables use: target grid -> source lon/lat)
    wl, wp = ka.rotated_to_geographic(lon, lat, o_lat_p * DEG, lon_0 * DEG)
    gl, gp = tr.values(proj, ll, lon, lat)
    arc = np.linalg.norm(ka._unit(gl, gp) - ka._unit(wl, wp), axis=-1)
    assert arc.max() <= 20 * bar, (tr.name, "inv", arc.max() / DEG)
    mid = (np.abs(wp) < 80 * DEG) & (np.abs(lat) < 80 * DEG)
    assert ka.angle_diff(gl[mid], wl[mid]).max() <= bar and np.abs(gp[mid] - wp[mid]).max() <= bar


def test_config2_target_mesh_against_change_of_basis(tr):
    """the very mesh bench.py regrids to (2000 x 2000, 0.0225 degree, rotated pole 22 N / lon_0 -40, every 7th point per axis):
    project_axes -> source lon/lat equals the 3-D rotation within 1e-9 degree of arc"""
    ax = ((np.arange(2000) - 999.5) * 0.0225)[::7] * DEG
    proj = "+proj=ob_tran +o_proj=longlat +lon_0=-40 +o_lat_p=22 +R=6.371e+06 +no_defs"
    ll = "+proj=latlong +a=6371000 +e=0 +no_defs"
    gl, gp = tr.axes(proj, ll, ax, ax)
    rl, rp = np.meshgrid(ax, ax)
    wl, wp = ka.rotated_to_geographic(rl.ravel(), rp.ravel(), 22 * DEG, -40 * DEG)
    arc = np.linalg.norm(ka._unit(gl, gp) - ka._unit(wl, wp), axis=-1)
    assert arc.max() <= ARC, arc.max() / DEG
    # the rotated origin is 68 N 40 W, the rotated pole 22 N 140 E (CF grid_north_pole_*)
    cl, cp = tr.values(proj, ll, [0.0, 0.0], [0.0, np.pi / 2])
    assert abs(cl[0] / DEG + 40) < 1e-9 and abs(cp[0] / DEG - 68) < 1e-9 and abs(cp[1] / DEG - 22) < 1e-9
    assert ka.angle_diff(cl[1], 140 * DEG) < 1e-9 * DEG
