"""The reference's CDMInterpolator tests that run on its own NetCDF-3 test files (test/testInterpolator.cc), restated through
the Interpolator mirror on the same data.  The arrays were extracted from the reference's files by
tests/golden/make_interpolator_fixtures.py (the GPU box has neither /root/reference nor a NetCDF reader); every test keeps
the reference's assertions and adds bit-parity with the oracle on the same inputs.

Not restated: test_interpolator / test_interpolatorKDTree / test_interpolatorRelative / test_interpolator_vectorlatlon need
the optional flth00.dat (hasTestExtra(), absent from the reference tree); test_interpolatorNcml and
test_interpolator_wrongaxes_latlon test NcML metadata.
test_interpolator_vector_backforth is in test_gpu_parity.py.
"""
import os

import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu

import fimex_b200 as fb  # noqa: E402
from fimex_b200 import Method  # noqa: E402

R = "6371000"
WGS84 = "+proj=latlong +datum=WGS84 +towgs84=0,0,0 +no_defs"


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "interpolator_fixtures.npz"))


def test_interpolator2coords(fx, oracle):
    # test/testInterpolator.cc:129-180: temp2(time, y_c, x_c) of twoCoordsTest.nc to a 12 x 12 polar-stereographic grid with
    # coord_kdtree and with nearestneighbor; more than 100 of the 144 cells of the first time step are above 29000
    proj = "+proj=stere +lat_0=90 +lon_0=0 +lat_ts=60 +ellps=sphere +a=" + R + " +e=0"
    x_axis = -1705516 + 50000.0 * np.arange(12)
    y_axis = -6872225 + 50000.0 * np.arange(12)
    temp2 = fx["two_temp2"][0]  # getDataSlice(varName) == unLimDimPos 0
    assert temp2.dtype == np.int16
    xc, yc = fx["two_x_c"].astype(np.float64), fx["two_y_c"].astype(np.float64)
    lon2, lat2 = fx["two_longitude2"].ravel(), fx["two_latitude2"].ravel()

    interp = fb.Interpolator(str(fx["two_proj4"]), xc, yc, False, lon2d=lon2, lat2d=lat2)
    interp.changeProjection(Method.COORD_NN_KD, proj, x_axis, y_axis, "m", "m")
    got = interp.getDataSlice(temp2)
    assert got.dtype == np.int16 and got.shape == (12, 12)
    assert int((got.astype(np.float64) > 29000).sum()) > 100
    # the same through the oracle: target lon/lat, kd search inside the distance of interest, gather, NaN -> fill, cast
    rc, tx, ty = oracle.project_axes(proj, WGS84, x_axis, y_axis)
    assert rc == 1
    dist = oracle.max_distance_of_interest(x_axis, y_axis, True)
    wx, wy, ties = oracle.coordkd(tx, ty, np.radians(lon2), np.radians(lat2), xc.size, yc.size, dist)
    assert ties == 0
    fill = fb.default_fill_value(np.int16)
    want = oracle.from_float(oracle.cached_interpolate(4, wx, wy, xc.size, yc.size, 12, 12, oracle.as_float(temp2, fill)[None]), fill, np.int16)
    assert np.array_equal(got, want.reshape(12, 12))

    interp = fb.Interpolator(str(fx["two_proj4"]), xc, yc, False, lon2d=lon2, lat2d=lat2)
    interp.changeProjection(Method.NEAREST_NEIGHBOR, proj, x_axis, y_axis, "m", "m")
    got = interp.getDataSlice(temp2)
    assert int((got.astype(np.float64) > 29000).sum()) > 100
    rc, px, py = oracle.project_axes(proj, str(fx["two_proj4"]), x_axis, y_axis)
    assert rc == 1
    px = oracle.points2position(px, xc, 0)
    py = oracle.points2position(py, yc, 0)
    want = oracle.from_float(oracle.cached_interpolate(0, px, py, xc.size, yc.size, 12, 12, oracle.as_float(temp2, fill)[None]), fill, np.int16)
    assert np.array_equal(got, want.reshape(12, 12))


def test_interpolatorSatellite(fx):
    # test/testInterpolator.cc:105-126: a swath with float 2-D lat/lon (fill -999 -> NaN by getScaledData) and no projection,
    # coord_kdtree to a 10 x 10 lat/lon grid.  The reference only checks that both getDataSlice overloads return the same
    # number of values (its axes put 55..55.9 on x and -106..-105.1 on y, i.e. swapped, so nothing is in range).
    lat = fx["sat_lat"].astype(np.float64)
    lon = fx["sat_lon"].astype(np.float64)
    lat[lat == -999.0] = np.nan
    lon[lon == -999.0] = np.nan
    cma = fx["sat_cma"][0]
    x_axis = 55 + 0.1 * np.arange(10)
    y_axis = -106 + 0.1 * np.arange(10)
    interp = fb.Interpolator("", np.arange(51.0), np.arange(51.0), False, lon2d=lon.ravel(), lat2d=lat.ravel())
    interp.changeProjection(Method.COORD_NN_KD, "+proj=latlon +R=" + R + " +e=0", x_axis, y_axis, "degrees_east", "degrees_north")
    got = interp.getDataSlice(cma)
    assert got.shape == (10, 10) and got.dtype == np.int16
    assert (got == fb.default_fill_value(np.int16)).all()
    # with the axes the right way round the swath is found: lon on x, lat on y
    interp.changeProjection(Method.COORD_NN_KD, "+proj=latlon +R=" + R + " +e=0", -106.5 + 0.1 * np.arange(10), 56.3 + 0.04 * np.arange(10),
                            "degrees_east", "degrees_north")
    got = interp.getDataSlice(cma)
    assert np.isin(got, (0, 1, -32767)).all() and np.isin(got, (0, 1)).sum() > 50


def _erai(fx):
    return (str(fx["erai_proj4"]), fx["erai_longitude"].astype(np.float64), fx["erai_latitude"].astype(np.float64), fx["erai_ga_skt"])


def _template_oracle(oracle, method, proj, lon, lat, tlon, tlat, ox, oy, field):
    rc, x, y = oracle.project_values(WGS84, proj, np.radians(tlon), np.radians(tlat))
    assert rc == 1
    py = oracle.points2position(y, np.radians(lat), 2)
    px = oracle.points2position(x, np.radians(lon), 1)
    f32 = oracle.as_float(field, fb.default_fill_value(np.float64))
    out = oracle.cached_interpolate(int(method), px, py, lon.size, lat.size, ox, oy, f32.reshape(-1, lat.size, lon.size))
    return oracle.from_float(out, fb.default_fill_value(np.float64), np.float64)


def test_interpolator_template(fx, oracle):
    # test/testInterpolator.cc:220-239: ga_skt (double, 0.75-degree lat/long, latitude descending) bicubic to the 2-D
    # longitude/latitude of template_noaa17.nc (29 x 31); only the first 7 data points are defined, between 270 and 280
    proj, lon, lat, skt = _erai(fx)
    tlon, tlat = fx["tmpl_longitude"].astype(np.float64), fx["tmpl_latitude"].astype(np.float64)
    assert tlon.shape == (31, 29)
    interp = fb.Interpolator(proj, lon, lat, True)
    interp.changeProjectionToTemplate(Method.BICUBIC, tlon, tlat)
    ci = interp.cachedInterpolation
    assert ci.getOutX() == 29 and ci.getOutY() == 31
    got = interp.getDataSlice(skt)  # getData("ga_skt"): all 8 times
    assert got.dtype == np.float64 and got.shape == (8, 1, 31, 29)
    fill = fb.default_fill_value(np.float64)
    arr = np.where(got == fill, np.nan, got).ravel()  # asDouble() after getScaledData: fill -> NaN
    assert (~np.isnan(arr[:7])).all() and (arr[:7] < 280).all() and (arr[:7] > 270).all()
    assert np.isnan(arr[7])
    want = _template_oracle(oracle, Method.BICUBIC, proj, lon, lat, tlon.ravel(), tlat.ravel(), 29, 31, skt)
    assert_bit_equal(got.ravel(), want.ravel(), "template bicubic (double in, double out)")
    with pytest.raises(fb.FimexB200Error):  # :706-714
        interp.changeProjectionToTemplate(Method.COORD_NN, tlon, tlat)


def test_interpolator_latlon(fx, oracle):
    # test/testInterpolator.cc:241-264: bilinear to a list of 10 geographic points
    lat_vals = [59.109, 59.052, 58.994, 58.934, 58.874, 58.812, 58.749, 58.685, 58.62, 64.0]
    lon_vals = [4.965, 5.13, 5.296, 5.465, 5.637, 5.81, 5.986, 6.164001, 6.344, 3.0]
    proj, lon, lat, skt = _erai(fx)
    interp = fb.Interpolator(proj, lon, lat, True)
    interp.changeProjectionToLonLatValues(Method.BILINEAR, lon_vals, lat_vals)
    ci = interp.cachedInterpolation
    assert ci.getOutX() == len(lon_vals) and ci.getOutY() == 1
    got = interp.getDataSlice(skt)
    assert got.shape == (8, 1, 1, 10)
    fill = fb.default_fill_value(np.float64)
    arr = np.where(got == fill, np.nan, got).ravel()
    assert not np.isnan(arr[0]) and 270 < arr[0] < 280
    assert (~np.isnan(arr)).all() and (arr < 281.1).all() and (arr > 266).all()
    # the values are narrowed to float first (double_to_float_cast, CDMInterpolator.cc:454-457)
    tlon = np.asarray(lon_vals, np.float32).astype(np.float64)
    tlat = np.asarray(lat_vals, np.float32).astype(np.float64)
    want = _template_oracle(oracle, Method.BILINEAR, proj, lon, lat, tlon, tlat, 10, 1, skt)
    assert_bit_equal(got.ravel(), want.ravel(), "lon/lat values bilinear")
    with pytest.raises(fb.FimexB200Error):
        interp.changeProjectionToLonLatValues(Method.BILINEAR, lon_vals, lat_vals[:-1])
    with pytest.raises(fb.FimexB200Error):
        interp.changeProjectionToLonLatValues(Method.FORWARD_MEAN, lon_vals, lat_vals)


def test_processor_rotate(fx, oracle):
    # test/testProcessor.cc:71-93: rotateAllVectorsToLatLon(true) on the 10 m wind of coordTest.nc; value 3 of each component
    # changes, the speed does not (1e-4 percent)
    proj = str(fx["coord_proj4"])
    x, y = fx["coord_x"].astype(np.float64), fx["coord_y"].astype(np.float64)
    xo, yo = fx["coord_x_wind_10m"], fx["coord_y_wind_10m"]
    proc = fb.Processor(proj, x, y, False).rotateVectorToLatLon(True)
    xn = proc.getDataSlice(xo, yo, "x")
    yn = proc.getDataSlice(yo, xo, "y")
    assert xn.dtype == np.float32 and xn.shape == xo.shape
    assert xn.ravel()[3] != xo.ravel()[3] and yn.ravel()[3] != yo.ravel()[3]
    a = float(xn.ravel()[3]) ** 2 + float(yn.ravel()[3]) ** 2
    b = float(xo.ravel()[3]) ** 2 + float(yo.ravel()[3]) ** 2
    assert abs(a - b) <= 1e-6 * max(a, b)
    np.testing.assert_allclose(xn.astype(np.float64) ** 2 + yn.astype(np.float64) ** 2, xo.astype(np.float64) ** 2 + yo.astype(np.float64) ** 2,
                               rtol=1e-5)
    # against the oracle: matrix of the expanded mesh (makeCachedVectorReprojection, CDMProcessor.cc:124-135), then rotation
    xf, yf = np.tile(x, y.size), np.repeat(y, x.size)
    rc, m = oracle.vector_matrix_field(proj, WGS84, xf, yf, x.size, y.size)
    assert rc == 1
    got_m = proc.cachedVectorReprojection.getMatrix()
    assert np.abs(got_m.reshape(-1, 4)[:, :3] - m.reshape(-1, 4)[:, :3]).max() < 1e-9
    u, v = oracle.vector_reproject_by_matrix(got_m, xo.copy().ravel()[None], yo.copy().ravel()[None], x.size, y.size, 1)
    assert_bit_equal(xn.ravel(), np.asarray(u).ravel(), "rotated x_wind_10m")
    assert_bit_equal(yn.ravel(), np.asarray(v).ravel(), "rotated y_wind_10m")
    # and back: geographic -> grid directions restores the wind
    back = fb.Processor(proj, x, y, False).rotateVectorToLatLon(False)
    xb, yb = back.getVectorSlices(xn, yn)
    np.testing.assert_allclose(xb, xo, atol=2e-3)
    np.testing.assert_allclose(yb, yo, atol=2e-3)
    rc, m2 = oracle.vector_matrix(WGS84, proj, x, y, 0, 0)
    assert rc == 1
    assert np.abs(back.cachedVectorReprojection.getMatrix().reshape(-1, 4)[:, :3] - m2.reshape(-1, 4)[:, :3]).max() < 1e-9


def test_processor_rotate_packed_short(oracle):
    # the rotation branch on a packed variable: fill -> NaN, rotate the raw values, NaN -> fill, round to short
    # (CDMProcessor.cc:604-616 with data2InterpolationArray / interpolationArray2Data)
    rot = "+proj=ob_tran +o_proj=longlat +lon_0=-40 +o_lat_p=22 +R=6.371e+06 +no_defs"
    x = -10 + 0.25 * np.arange(80)
    y = -8 + 0.25 * np.arange(60)
    rng = np.random.default_rng(5)
    u = rng.integers(-3000, 3000, (3, y.size, x.size)).astype(np.int16)
    v = rng.integers(-3000, 3000, (3, y.size, x.size)).astype(np.int16)
    u[rng.random(u.shape) < 0.05] = -32767
    v[rng.random(v.shape) < 0.05] = -32767
    proc = fb.Processor(rot, x, y, True).rotateVectorToLatLon(True)
    un, vn = proc.getVectorSlices(u, v)
    assert un.dtype == np.int16 and un.shape == u.shape
    m = proc.cachedVectorReprojection.getMatrix()
    fu, fv = oracle.as_float(u, -32767), oracle.as_float(v, -32767)
    ru, rv = oracle.vector_reproject_by_matrix(m, fu.reshape(3, -1), fv.reshape(3, -1), x.size, y.size, 3)
    assert np.array_equal(un.ravel(), oracle.from_float(np.asarray(ru), -32767, np.int16).ravel())
    assert np.array_equal(vn.ravel(), oracle.from_float(np.asarray(rv), -32767, np.int16).ravel())
    mask = (u == -32767) | (v == -32767)
    assert (un[mask] == -32767).all() and (vn[mask] == -32767).all()
    # directions: same matrix through reprojectDirectionValues
    ang = rng.uniform(0, 360, (2, y.size, x.size)).astype(np.float32)
    got = proc.getDirectionSlice(ang)
    want = oracle.vector_reproject_direction(m, ang.reshape(2, -1).copy(), x.size, y.size, 2)
    assert_bit_equal(got.ravel(), np.asarray(want).ravel(), "rotated directions")
    with pytest.raises(fb.FimexB200Error):
        fb.Processor(rot, x, y, True).getVectorSlices(u, v)


def test_interpolator_vcross(fx, oracle):
    # test/testInterpolator.cc:474-497: two cross-sections over the 0.75-degree ERA-Interim field, bilinear
    proj, lon, lat, skt = _erai(fx)
    vc = [("OsloTrondheimTromso", [(10.74, 59.9), (10.3951, 63.4305), (18.9551, 69.6489)]),
          ("BergenOslo", [(5.3290, 60.3983), (10.74, 59.9)])]
    interp = fb.Interpolator(proj, lon, lat, True)
    interp.changeProjectionToCrossSections(Method.BILINEAR, vc)
    assert interp.vcross_names == ["OsloTrondheimTromso", "BergenOslo"] and len(interp.vcross_bnds) == 2  # nvcross == 2
    ci = interp.cachedInterpolation
    assert ci.getOutX() > 5 and ci.getOutY() == 1
    # restated: legs sampled in the source CRS (degrees here: a lat/long grid), floor(max(|dx/0.75|, |dy/0.75|)) steps per leg
    want_lon, want_lat, starts = [], [], []
    for _, pts in vc:
        starts.append(len(want_lon))
        for i in range(1, len(pts)):
            (x0, y0), (x1, y1) = pts[i - 1], pts[i]
            num = int(np.floor(max(abs((x1 - x0) / (lon[1] - lon[0])), abs((y1 - y0) / (lat[1] - lat[0])))))
            if i == 1:
                want_lon.append(x0), want_lat.append(y0)
            for j in range(1, num):
                want_lon.append(x0 + j * (x1 - x0) / num), want_lat.append(y0 + j * (y1 - y0) / num)
            want_lon.append(x1), want_lat.append(y1)
    assert ci.getOutX() == len(want_lon)
    assert interp.vcross_bnds == [(starts[0], starts[1] - 1), (starts[1], len(want_lon) - 1)]
    # the source grid is itself lat/long on the same sphere, so the round trip through the projection is (nearly) the identity
    tlon = np.asarray(want_lon, np.float32).astype(np.float64)
    tlat = np.asarray(want_lat, np.float32).astype(np.float64)
    got = interp.getDataSlice(skt)
    assert got.shape == (8, 1, 1, len(want_lon))
    want = _template_oracle(oracle, Method.BILINEAR, proj, lon, lat, tlon, tlat, len(want_lon), 1, skt)
    fill = fb.default_fill_value(np.float64)
    defined = want.ravel() != fill
    assert defined.sum() >= 8 * 3  # only the Bergen end of the second section lies inside the 3..6.75 E x 57..64.5 N field
    assert np.array_equal(got.ravel() != fill, defined)
    np.testing.assert_allclose(got.ravel()[defined], want.ravel()[defined], rtol=1e-6)
    # a projected source: the legs are straight in the projection plane, one point per 50 km cell
    stere = "+proj=stere +lat_0=90 +lon_0=0 +lat_ts=60 +units=m +a=6.371e+06 +e=0 +no_defs"
    x = -1705516 + 50162.0 * np.arange(40)
    y = -6872225 + 50162.0 * np.arange(40)
    ip = fb.Interpolator(stere, x, y, False)
    ip.changeProjectionToCrossSections("bilinear", [("a", [(-13.0, 32.0), (-8.0, 44.0)]), ("p", [(-10.0, 40.0)])])
    assert ip.vcross_bnds[1][0] == ip.vcross_bnds[1][1] == ip.cachedInterpolation.getOutX() - 1
    gx, gy = ip.cachedInterpolation.points()
    n = ip.vcross_bnds[0][1] + 1
    assert n > 10 and np.abs(np.diff(gx[:n], 2)).max() < 1e-3 and np.abs(np.diff(gy[:n], 2)).max() < 1e-3  # equidistant in index space
    assert np.hypot(np.diff(gx[:n]), np.diff(gy[:n])).max() <= 1.5


def test_interpolator_relative_axes(oracle):
    # test_interpolatorRelative (test/testInterpolator.cc:182-196, needs the optional flth00.dat there): axes given relative to
    # the bounding box of the data in the target projection, "0,50000,...,x;relativeStart=0" (CDMInterpolator.cc:345-412)
    lon = np.arange(-10.0, 30.25, 0.5)
    lat = np.arange(50.0, 72.25, 0.5)
    proj = "+proj=stere +lat_0=90 +lon_0=-32 +lat_ts=60 +ellps=sphere +a=6371000 +e=0"
    interp = fb.Interpolator("+proj=latlong +a=6371000 +e=0 +no_defs", lon, lat, True)
    interp.changeProjection("bilinear", proj, "0,50000,...,x;relativeStart=0", "0,50000,...,x;relativeStart=0", "m", "m")
    ci = interp.cachedInterpolation
    lon2d, lat2d = np.tile(lon, lat.size), np.repeat(lat, lon.size)
    rc, x, y = oracle.project_values(WGS84, proj, np.radians(lon2d), np.radians(lat2d))
    assert rc == 1
    want_x = fb.spatial_axis_spec("0,50000,...,x;relativeStart=0", x.min(), x.max())
    want_y = fb.spatial_axis_spec("0,50000,...,x;relativeStart=0", y.min(), y.max())
    assert ci.getOutX() == want_x.size and ci.getOutY() == want_y.size
    assert want_x[0] >= x.min() - 50000 and want_x[-1] <= x.max() and np.all(np.diff(want_x) == 50000)
    field = np.random.default_rng(3).normal(280, 5, (2, lat.size, lon.size)).astype(np.float32)
    got = interp.getDataSlice(field)
    assert got.shape == (2, want_y.size, want_x.size)
    inside = np.isfinite(got) & (got != np.float32(fb.default_fill_value(np.float32)))
    assert inside.mean() > 0.5  # the box of a lat/lon rectangle in a polar projection has empty corners
    # same table as with the explicit axes
    ref = fb.Interpolator("+proj=latlong +a=6371000 +e=0 +no_defs", lon, lat, True)
    ref.changeProjection("bilinear", proj, want_x, want_y, "m", "m")
    assert_bit_equal(got, ref.getDataSlice(field), "relative axes")
    with pytest.raises(fb.FimexB200Error):  # "only implemented for projections in m, not degree yet"
        interp.changeProjection("bilinear", "+proj=latlong +R=6371000", "0,1,...,x;relativeStart=0", "0,1,...,x;relativeStart=0", "degree", "degree")


def test_config1_hirlam12_real_file(fx, oracle):
    """BASELINE config 1 on the reference's OWN file: test/hirlam12.nc (time=2, pressure=2, Yc=12, Xc=17, float32 with
    _FillValue 9.96921e+36, axes in degrees) through the CLI's option strings --interpolate.method=bilinear
    --interpolate.projString="+proj=latlong ..." --interpolate.xAxisValues=5,5.5,6,6.5 --interpolate.yAxisValues=61.5,62,62.5
    --interpolate.{x,y}AxisUnit=degree (src/binSrc/fimex.cc:1008-1034).  Golden = the COMPILED reference's mifi_interpolate_f /
    mifi_vector_reproject_values_f on the same fields (tests/golden/make_interpolator_fixtures.py); bit for bit, fill values
    where the reference has NaN (interpolationArray2Data, CDMInterpolator.cc:121-124)."""
    sphere = "+proj=latlong +a=6371000 +e=0 +no_defs"
    fill = np.float32(fx["hirlam_fill"])
    assert fill == np.float32(9.96921e+36)
    ip = fb.Interpolator(sphere, fx["hirlam_Xc"], fx["hirlam_Yc"], True, has_xy_vectors=True)
    ip.changeProjection("bilinear", sphere, "5,5.5,6,6.5", "61.5,62,62.5", "degree", "degree")
    for name in ("geopotential_height", "air_potential_temperature"):
        data = fx["hirlam_" + name]
        assert data.shape == (2, 2, 12, 17) and (data == fill).sum() > 0  # the file has undefined rows
        golden = fx["hirlam_golden_" + name]
        for t in range(2):  # one getDataSlice per unlimited-dimension position, as the writer calls it
            got = ip.getDataSlice(data[t], bad_value=float(fill))
            want = golden[2 * t:2 * t + 2]
            assert got.dtype == np.float32 and got.shape == (2, 3, 4)
            assert np.array_equal(got == fill, np.isnan(want)), name
            assert_bit_equal(np.where(got == fill, np.float32(np.nan), got), want, f"hirlam12 {name} t={t}")
        assert np.isnan(golden).sum() > 0 and (~np.isnan(golden)).sum() > 0
    # x_wind / y_wind: spatial vectors, rotated with the lat/long-target bearing branch (interpolation.c:366,408); each
    # component's call interpolates both and rotates (CDMInterpolator.cc:260-283)
    xw, yw = fx["hirlam_x_wind"], fx["hirlam_y_wind"]
    for t in range(2):
        gu = ip.getDataSlice(xw[t], counterpart=yw[t], direction="x", bad_value=float(fill))
        gv = ip.getDataSlice(yw[t], counterpart=xw[t], direction="y", bad_value=float(fill))
        for got, key in ((gu, "x_wind_rotated"), (gv, "y_wind_rotated")):
            want = fx["hirlam_golden_" + key][2 * t:2 * t + 2]
            assert np.array_equal(got == fill, np.isnan(want)), key
            ok = ~np.isnan(want)
            assert np.abs(got[ok] - want[ok]).max() <= 1e-5 * np.abs(want[ok]).max(), key  # north_star: <= 1e-5 relative for vector output
    # and the same numbers through the library's own one-shot C symbol (the reference's prototype)
    f = fx["hirlam_air_potential_temperature"].reshape(4, 12, 17).copy()
    f[f == fill] = np.nan
    rc, out = fb.mifi_interpolate_f(Method.BILINEAR, sphere, f, fx["hirlam_Xc"], fx["hirlam_Yc"], fb.LONGITUDE, fb.LATITUDE, 4, sphere,
                                    fx["hirlam_target_x"], fx["hirlam_target_y"], fb.LONGITUDE, fb.LATITUDE)
    assert rc == fb.MIFI_OK
    assert_bit_equal(out.reshape(4, 3, 4), fx["hirlam_golden_air_potential_temperature"], "hirlam12 mifi_interpolate_f")


def test_merger(fx, oracle):
    """test/testMerger.cc:44-77 (test_merger) on the reference's own files: ga_2t_1 of test_merge_inner.nc merged into
    test_merge_outer.nc on the inner grid extended over the outer one (61 x 113), bilinear, linear border smoothing.  The
    reference's three known answers -- middle, transition zone, outer -- within its own tolerance 1e-3, and the whole merged
    field against the same chain evaluated with the CPU oracle's kernels."""
    proj = str(fx["merge_inner_proj4"])
    assert proj == str(fx["merge_outer_proj4"])
    xi, yi = fx["merge_inner_longitude"].astype(np.float64), fx["merge_inner_latitude"].astype(np.float64)
    xo, yo = fx["merge_outer_longitude"].astype(np.float64), fx["merge_outer_latitude"].astype(np.float64)
    vi, vo = fx["merge_inner_ga_2t_1"][0, 0], fx["merge_outer_ga_2t_1"][0, 0]
    merger = fb.Merger(proj, xi, yi, proj, xo, yo, True)
    merger.setTargetGridFromInner()
    got = merger.getDataSlice(vi, vo)
    NLON, NLAT = 61, 113
    assert got.shape == (NLAT, NLON) and got.dtype == np.float64
    for ilon, ilat, expected in ((28, 56, 288.104), (24, 56, 288.467), (8, 56, 289.937)):
        assert abs(got[ilat, ilon] - expected) < 0.001, (ilon, ilat, got[ilat, ilon])

    def interp(src_x, src_y, field, tx, ty):
        rc, x, y = oracle.project_axes(proj, proj, np.radians(tx), np.radians(ty))
        px = oracle.points2position(x, np.radians(src_x), 1)
        py = oracle.points2position(y, np.radians(src_y), 2)
        return oracle.cached_interpolate(1, px, py, src_x.size, src_y.size, tx.size, ty.size, field.astype(np.float32)[None])[0].astype(np.float64)

    smooth = fb.linear_border_smoothing(vi, interp(xo, yo, vo, xi, yi))
    top, base = interp(xi, yi, smooth, merger.target_x, merger.target_y), interp(xo, yo, vo, merger.target_x, merger.target_y)
    want = np.where(np.isnan(top), base, top)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
