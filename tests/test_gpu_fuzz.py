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
