"""Host logic of the CDMMerger mirror (fimex_b200/merger.py): axis extension and the linear border blend -- no GPU needed.
Reference: src/CDMMerger.cc:233-274, src/CDMBorderSmoothing_Linear.cc:43-84; numbers of test/testMerger.cc:44-77."""
import os

import numpy as np
import pytest

from fimex_b200.merger import extend_inner_axis, linear_border_smoothing
from fimex_b200.capi import FimexB200Error


def test_extend_inner_axis_reference_files():
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "interpolator_fixtures.npz"))
    vx = extend_inner_axis(fx["merge_inner_longitude"], fx["merge_outer_longitude"])
    vy = extend_inner_axis(fx["merge_inner_latitude"], fx["merge_outer_latitude"])
    assert (vx.size, vy.size) == (61, 113)  # NLON, NLAT of the reference's test_merger
    assert vx[0] == -35.0 and vx[-1] == -20.0 and np.allclose(np.diff(vx), 0.25)
    assert vy[0] == 54.5 and vy[-1] == 40.5 and np.allclose(np.diff(vy), -0.125)  # descending axes stay descending


def test_extend_inner_axis_errors():
    with pytest.raises(FimexB200Error):
        extend_inner_axis([0.0, 1.0, 2.5], [-5.0, 5.0])  # not equidistant
    with pytest.raises(FimexB200Error):
        extend_inner_axis([0.0, 1.0, 2.0], [0.5, 5.0])  # inner not inside outer
    with pytest.raises(FimexB200Error):
        extend_inner_axis([1.0], [0.0, 5.0])


def test_linear_border_smoothing():
    ny, nx = 30, 40
    inner, outer = np.full((ny, nx), 10.0), np.full((ny, nx), 20.0)
    m = linear_border_smoothing(inner, outer)  # transition 5, border 2
    assert (m[:2] == 20).all() and (m[-2:] == 20).all() and (m[:, :2] == 20).all() and (m[:, -2:] == 20).all()
    assert (m[7:ny - 7, 7:nx - 7] == 10).all()
    # the ramp along the left edge, away from the corners: alpha = (xmax1 - x) / 5 for x = 2..6
    assert np.allclose(m[15, 2:7], 10 + 10 * np.array([5, 4, 3, 2, 1]) / 5)
    # corner: Euclidean distance, clipped to 1
    assert m[3, 3] == 20 and np.isclose(m[6, 6], 10 + 10 * np.sqrt(2) / 5)
    inner[15, 20] = np.nan
    outer[10, 10] = np.nan
    m = linear_border_smoothing(inner, outer)
    assert m[15, 20] == 20 and m[10, 10] == 10  # inner undefined -> outer; outer undefined -> inner
    assert np.isnan(linear_border_smoothing(inner, outer, use_outer_if_inner_undefined=False)[15, 20])
