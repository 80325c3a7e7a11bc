"""BASELINE.json configs 3, 4 and 5 at FULL size on the GPU (config 2 is covered by test_gpu_parity.py::test_full_size_*
and by bench.py): size-independent properties plus bit-for-bit comparison against the oracle on samples the CPU can
finish in seconds.  Timings are appended to gpurun_out/config_timings.jsonl (informational; bench.py is the metric).
"""
import json
import os
import time

import numpy as np
import pytest

from conftest import ROOT, assert_bit_equal

pytestmark = pytest.mark.gpu

import fimex_b200 as fb  # noqa: E402
from fimex_b200 import Method  # noqa: E402

SRC_LL = "+proj=latlong +a=6371000 +e=0 +no_defs"
ROTPOLE = "+proj=ob_tran +o_proj=longlat +lon_0=-40 +o_lat_p=22 +R=6.371e+06 +no_defs"
STERE = "+proj=stere +lat_0=90 +lon_0=0 +lat_ts=60 +a=6371000 +e=0"
LCC = "+proj=lcc +lat_0=63 +lon_0=15 +lat_1=63 +lat_2=63 +no_defs +R=6.371e+06"
WGS84 = "+proj=latlong +datum=WGS84 +towgs84=0,0,0 +no_defs"
LON = np.arange(1440) * 0.25
LAT = 90 - np.arange(721) * 0.25
NZ = 137


def _log(name, **kw):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "config_timings.jsonl"), "a") as f:
            f.write(json.dumps({"config": name, **kw}) + "\n")


def _time_device(fn, reps=3):
    import torch
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def test_config3_nearest_neighbour_full_size(oracle):
    """config 3 (i): nearestneighbor through the projection path, 2000x2000 x 137 levels: index table equal to the
    oracle's, output memcmp-equal including the NaN pattern 0x7fc00000"""
    import torch
    ax = (np.arange(2000) - 999.5) * 0.0225
    t0 = time.perf_counter()
    ci = fb.CachedInterpolation.fromProjection(Method.NEAREST_NEIGHBOR, ROTPOLE, ax, ax, True, True, SRC_LL, LON, LAT, True)
    assert ci.createReducedDomain()
    setup = time.perf_counter() - t0
    inX, inY = ci.getInX(), ci.getInY()
    x0, y0 = ci.reducedDomain()[2:4]
    gx, gy = ci.points()
    rc, x, y = oracle.project_axes(ROTPOLE, SRC_LL, np.radians(ax), np.radians(ax))
    assert rc == 1
    ox = oracle.points2position(x, np.radians(LON), 1) - x0
    oy = oracle.points2position(y, np.radians(LAT), 2) - y0
    assert np.abs(gx - ox).max() <= 4e-9 and np.abs(gy - oy).max() <= 4e-9  # 1e-9 degree in 0.25-degree cells
    lr = lambda v: np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5))  # lround: half away from zero
    flips = int(((lr(gx) != lr(ox)) | (lr(gy) != lr(oy))).sum())
    near = int(((np.abs(gx - np.floor(gx) - 0.5) < 1e-9) | (np.abs(gy - np.floor(gy) - 0.5) < 1e-9)).sum())
    assert flips <= max(2, near), (flips, near)  # index table identical except points within 1e-9 of a half-cell boundary
    g = torch.Generator(device="cuda").manual_seed(20261018)
    field = torch.randn((NZ, inY, inX), generator=g, device="cuda", dtype=torch.float32)
    field[torch.rand(field.shape, generator=g, device="cuda") < 0.01] = float("nan")
    out = ci.interpolateValues(field)
    ms = _time_device(lambda: ci.interpolateValues(field, out=out.view(-1)))
    rng = np.random.default_rng(3)
    sample = rng.integers(0, 4_000_000, 3000)
    want = oracle.cached_interpolate(0, gx[sample], gy[sample], inX, inY, sample.size, 1, field.cpu().numpy())
    got = out.view(NZ, -1)[:, torch.from_numpy(sample).cuda()].cpu().numpy().reshape(want.shape)
    assert_bit_equal(got, want, "config 3 NN sample", nan_payload=True)
    # idempotence: regridding the field onto its own grid with NN returns the field (copied values, bit for bit)
    ident = fb.CachedInterpolation("x", "y", Method.NEAREST_NEIGHBOR, np.tile(np.arange(inX, dtype=float), inY),
                                   np.repeat(np.arange(inY, dtype=float), inX), inX, inY, inX, inY)
    same = ident.interpolateValues(field[:4].contiguous())
    assert torch.equal(same.view(torch.int32), field[:4].view(torch.int32))
    _log("3i nearestneighbor 2000x2000x137", ms=ms, values_per_s=NZ * 4e6 / (ms * 1e-3), setup_s=setup, index_flips=flips, near_half_cell=near)


def test_config3_coord_nearestneighbor_full_size(oracle):
    """config 3 (ii): coord_nearestneighbor, 2-D lon/lat of the 1440x721 source, 2000x2000 targets, no crop"""
    import torch
    ax = (np.arange(2000) - 999.5) * 0.0225
    lon2d, lat2d = oracle.lonlat_to_matrix(np.radians(LON), np.radians(LAT))
    t0 = time.perf_counter()
    ci = fb.CachedInterpolation.fromCoordinates(Method.COORD_NN, ROTPOLE, ax, ax, True, True, np.degrees(lon2d), np.degrees(lat2d), 1440, 721)
    setup = time.perf_counter() - t0
    assert (ci.getInX(), ci.getInY()) == (1440, 721)  # not cropped on this path
    gx, gy = ci.points()
    assert ((gx >= 0) & (gx < 1440) & (gy >= 0) & (gy < 721)).all()
    # oracle search for a random subset of targets
    rc, tx, ty = oracle.project_axes(ROTPOLE, WGS84, np.radians(ax), np.radians(ax))
    rng = np.random.default_rng(5)
    sample = rng.integers(0, 4_000_000, 4000)
    wx, wy, ties = oracle.coordnn(tx[sample], ty[sample], lon2d, lat2d, 1440, 721)
    differ = int(((gx[sample] != wx) | (gy[sample] != wy)).sum())
    assert differ <= max(3, ties), (differ, ties)
    field = torch.randn((NZ, 721, 1440), device="cuda", dtype=torch.float32)
    out = ci.interpolateValues(field)
    ms = _time_device(lambda: ci.interpolateValues(field, out=out.view(-1)))
    want = oracle.cached_interpolate(3, gx[sample], gy[sample], 1440, 721, sample.size, 1, field[:8].cpu().numpy())
    got = out.view(NZ, -1)[:8, torch.from_numpy(sample).cuda()].cpu().numpy().reshape(want.shape)
    assert_bit_equal(got, want, "config 3 coord_nn sample", nan_payload=True)
    _log("3ii coord_nearestneighbor 2000x2000x137", ms=ms, values_per_s=NZ * 4e6 / (ms * 1e-3), setup_s=setup, sample_mismatch=differ, oracle_ties=int(ties))


def test_config4_bicubic_vector_full_size(oracle):
    """config 4: bicubic regrid of x_wind / y_wind (137 levels) with CachedVectorReprojection to a polar-stereographic
    3000x3000 grid; NaN wedge along the 0/360 seam and north of 89.75 degrees (no wrap, no edge fallback)"""
    import torch
    ax = -3748750.0 + 2500.0 * np.arange(3000)
    t0 = time.perf_counter()
    ci = fb.CachedInterpolation.fromProjection(Method.BICUBIC, STERE, ax, ax, False, False, SRC_LL, LON, LAT, True)
    assert ci.createReducedDomain()
    cvr = fb.CachedVectorReprojection.fromProjection(fb.MIFI_VECTOR_KEEP_SIZE, SRC_LL, STERE, ax, ax, fb.PROJ_AXIS, fb.PROJ_AXIS)
    setup = time.perf_counter() - t0
    inX, inY = ci.getInX(), ci.getInY()
    g = torch.Generator(device="cuda").manual_seed(4)
    u = torch.randn((NZ, inY, inX), generator=g, device="cuda", dtype=torch.float32) * 10
    v = torch.randn((NZ, inY, inX), generator=g, device="cuda", dtype=torch.float32) * 10
    uo, vo = ci.interpolateVector(u, v, cvr)
    ms = _time_device(lambda: ci.interpolateVector(u, v, cvr), reps=2)
    n = 9_000_000
    nan_u = torch.isnan(uo[0])
    assert torch.equal(nan_u, torch.isnan(vo[0])) and torch.equal(nan_u, torch.isnan(uo[-1]))
    frac_nan = nan_u.float().mean().item()
    assert 0.0 < frac_nan < 0.05  # the seam wedge + the polar cap, a thin part of the domain
    gx, gy = ci.points()
    valid = (np.floor(gx) >= 1) & (np.floor(gx) + 2 < inX) & (np.floor(gy) >= 1) & (np.floor(gy) + 2 < inY)
    assert np.array_equal(~valid.reshape(3000, 3000), nan_u.cpu().numpy())  # exactly the reference's validity rule (:975-976)
    # the rotation preserves the vector length (MIFI_VECTOR_KEEP_SIZE): compare with the un-rotated interpolation
    pu, pv = ci.interpolateVector(u[:2].contiguous(), v[:2].contiguous(), None)
    l0 = torch.sqrt(pu.double()**2 + pv.double()**2)
    l1 = torch.sqrt(uo[:2].double()**2 + vo[:2].double()**2)
    ok = ~torch.isnan(l0)
    assert ((l0 - l1).abs()[ok] <= 1e-5 * (1 + l0[ok])).all()
    # sample against the oracle, bit for bit
    rng = np.random.default_rng(9)
    sample = rng.integers(0, n, 1500)
    m = cvr.getMatrix().reshape(n, 4)[sample].ravel()
    hu, hv = u[:16].cpu().numpy(), v[:16].cpu().numpy()
    wu = oracle.cached_interpolate(2, gx[sample], gy[sample], inX, inY, sample.size, 1, hu)
    wv = oracle.cached_interpolate(2, gx[sample], gy[sample], inX, inY, sample.size, 1, hv)
    wu, wv = oracle.vector_reproject_by_matrix(m, wu, wv, sample.size, 1, 16)
    idx = torch.from_numpy(sample).cuda()
    assert_bit_equal(uo.view(NZ, -1)[:16, idx].cpu().numpy().ravel(), wu.ravel(), "config 4 u sample")
    assert_bit_equal(vo.view(NZ, -1)[:16, idx].cpu().numpy().ravel(), wv.ravel(), "config 4 v sample")
    _log("4 bicubic u/v + rotation 3000x3000x137", ms=ms, pairs_per_s=NZ * n / (ms * 1e-3), setup_s=setup, nan_fraction=frac_nan)


def _swath(oracle, ny=5000, nx=2000, seed=20261020):
    """synthetic sun-synchronous swath crossing 63N 15E: 5000 scan lines x 2000 pixels at 1 km
    spacing.  The ground-track azimuth at 63N for an inclination of 98.7 degrees follows from cos(i) = sin(az) cos(lat).
    The lattice is laid out in the target plane and taken back to longitude/latitude with the oracle, then jittered, so
    the input looks like what a swath file holds: 2-D longitude(y,x) / latitude(y,x) in degrees."""
    rng = np.random.default_rng(seed)
    az = np.arcsin(np.cos(np.radians(98.7)) / np.cos(np.radians(63.0)))  # about -19.5 degrees
    along = (np.arange(ny) - ny / 2)[:, None] * 1000.0
    across = (np.arange(nx) - nx / 2)[None, :] * 1000.0
    x = along * np.sin(az) + across * np.cos(az)
    y = along * np.cos(az) - across * np.sin(az)
    rc, lon, lat = oracle.project_values(LCC, WGS84, x.ravel(), y.ravel())
    assert rc == 1
    lon = np.degrees(lon).reshape(ny, nx) + rng.normal(0, 2e-3, (ny, nx))
    lat = np.degrees(lat).reshape(ny, nx) + rng.normal(0, 1e-3, (ny, nx))
    val = (280 + 10 * np.sin(np.radians(lat) * 20) + rng.normal(0, 0.5, lat.shape)).astype(np.float32)
    val[rng.random(val.shape) < 0.02] = np.nan
    return lon, lat, val


@pytest.mark.parametrize("method", [Method.FORWARD_MEAN, Method.FORWARD_MAX])
def test_config5_swath_forward_full_size(oracle, method):
    """config 5: 10 M swath points, forward_mean / forward_max onto a 1 km Lambert grid 3000 x 6000"""
    import torch
    lon, lat, val = _swath(oracle)
    ny, nx = lon.shape
    ox = -1500e3 + 1000.0 * np.arange(3000)
    oy = -3000e3 + 1000.0 * np.arange(6000)
    t0 = time.perf_counter()
    cfi = fb.CachedForwardInterpolation.fromCoordinates(method, LCC, ox, oy, False, False, lon.ravel(), lat.ravel(), nx, ny)
    setup = time.perf_counter() - t0
    gx, gy = cfi.points()
    # positions against the oracle pipeline (CDMInterpolator.cc:1265-1317) on a subset
    rng = np.random.default_rng(1)
    sub = rng.integers(0, nx * ny, 200_000)
    rc, x, y = oracle.project_values(WGS84, LCC, np.radians(lon.ravel()[sub]), np.radians(lat.ravel()[sub]))
    assert rc == 1
    x = oracle.points2position(x, ox, 0)
    y = oracle.points2position(y, oy, 0)
    assert np.abs(gx[sub] - x).max() < 1e-6 and np.abs(gy[sub] - y).max() < 1e-6
    inside = ((gx > -0.5) & (gx < 2999.5) & (gy > -0.5) & (gy < 5999.5)).mean()
    assert inside > 0.85  # the 3000 x 6000 km grid encloses the swath except the corners of the tilted strip
    d_val = torch.from_numpy(val).cuda().view(1, ny, nx).contiguous()
    out = cfi.interpolateValues(d_val)
    ms = _time_device(lambda: cfi.interpolateValues(d_val, out=out.view(-1)))
    want = oracle.forward_interpolate(int(method), gx, gy, nx, ny, 3000, 6000, val)
    assert_bit_equal(out.cpu().numpy(), want, f"config 5 {method.name}")
    filled = (~np.isnan(want)).mean()
    assert 0.25 < filled < 0.9
    n_in, n_cells = nx * ny, 3000 * 6000
    alg = 8 * n_in + 8 * n_cells  # values + perm + offsets + store
    _log(f"5 {method.name.lower()} 10M swath -> 3000x6000", ms=ms, points_per_s=n_in / (ms * 1e-3), gbs=alg / (ms * 1e-3) / 1e9, setup_s=setup,
         filled_fraction=float(filled))


def test_config2_get_data_slice_and_vector_full_size(oracle):
    """config 2 geometry (2000 x 2000 rotated pole, 137 levels) through the fused slice calls: a packed int16 variable in and
    out (fb200_interp_get_data_slice) and an x/y wind pair with rotation through the staged bilinear gather
    (fb200_interp_interpolate_vector); samples against the oracle, exact"""
    import torch
    ax = (np.arange(2000) - 999.5) * 0.0225
    ci = fb.CachedInterpolation.fromProjection(Method.BILINEAR, ROTPOLE, ax, ax, True, True, SRC_LL, LON, LAT, True)
    assert ci.createReducedDomain()
    cvr = fb.CachedVectorReprojection.fromProjection(fb.MIFI_VECTOR_KEEP_SIZE, SRC_LL, ROTPOLE, ax, ax, fb.LONGITUDE, fb.LATITUDE)
    inX, inY = ci.getInX(), ci.getInY()
    g = torch.Generator(device="cuda").manual_seed(7)
    packed = torch.randint(-3000, 3000, (NZ, inY, inX), generator=g, device="cuda", dtype=torch.int16)
    packed[torch.rand((NZ, inY, inX), generator=g, device="cuda") < 0.01] = -32767
    out = ci.getDataSlice(packed, -32767.0)
    ms16 = _time_device(lambda: ci.getDataSlice(packed, -32767.0, out=out.view(-1)))
    assert out.dtype == torch.int16 and out.shape == (NZ, 2000, 2000)
    frac_fill = (out[0] == -32767).float().mean().item()
    assert 0.02 < frac_fill < 0.06  # about 4 taps x 1 % of the points see an undefined tap
    gx, gy = ci.points()
    rng = np.random.default_rng(3)
    sample = rng.integers(0, 4_000_000, 3000)
    hp = packed[:8].cpu().numpy()
    want = oracle.from_float(oracle.cached_interpolate(1, gx[sample], gy[sample], inX, inY, sample.size, 1, oracle.as_float(hp, -32767.0)),
                             -32767.0, np.int16)
    idx = torch.from_numpy(sample).cuda()
    assert np.array_equal(out.view(NZ, -1)[:8, idx].cpu().numpy().ravel(), want.ravel())
    # x/y wind pair
    u = torch.randn((NZ, inY, inX), generator=g, device="cuda") * 10
    v = torch.randn((NZ, inY, inX), generator=g, device="cuda") * 10
    uo, vo = ci.interpolateVector(u, v, cvr)
    msv = _time_device(lambda: ci.interpolateVector(u, v, cvr), reps=2)
    m = cvr.getMatrix().reshape(-1, 4)[sample].ravel()
    hu, hv = u[:8].cpu().numpy(), v[:8].cpu().numpy()
    wu = oracle.cached_interpolate(1, gx[sample], gy[sample], inX, inY, sample.size, 1, hu)
    wv = oracle.cached_interpolate(1, gx[sample], gy[sample], inX, inY, sample.size, 1, hv)
    wu, wv = oracle.vector_reproject_by_matrix(m, wu, wv, sample.size, 1, 8)
    assert_bit_equal(uo.view(NZ, -1)[:8, idx].cpu().numpy().ravel(), wu.ravel(), "config 2 u sample")
    assert_bit_equal(vo.view(NZ, -1)[:8, idx].cpu().numpy().ravel(), wv.ravel(), "config 2 v sample")
    _log("2 getDataSlice int16 2000x2000x137", ms=ms16, values_per_s=NZ * 4e6 / (ms16 * 1e-3), fill_fraction=frac_fill)
    _log("2 bilinear u/v + rotation 2000x2000x137", ms=msv, pairs_per_s=NZ * 4e6 / (msv * 1e-3))
