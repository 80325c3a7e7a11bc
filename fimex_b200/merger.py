"""Host-side mirror of CDMMerger -- a "free beneficiary" of the regridding path: every horizontal
interpolation it does goes through the same tables and gather kernels.

Reference (arebru/fimex 0.67.2):
  CDMMerger::setTargetGridFromInner / setTargetGrid / getDataSlice     src/CDMMerger.cc:153-207
  CDMMergerPrivate::makeCDM (smooth -> interpolate -> overlay)          src/CDMMerger.cc:211-229
  CDMMergerPrivate::extendInnerAxis                                     src/CDMMerger.cc:233-274
  CDMBorderSmoothing::getDataSlice (inner/outer blend on the inner grid) src/CDMBorderSmoothing.cc:97-158
  CDMBorderSmoothing_Linear::operator()                                 src/CDMBorderSmoothing_Linear.cc:43-84
  CDMOverlay::getDataSlice (top where defined, else base)               src/CDMOverlay.cc:63-92
  makeMergedCDM (outer interpolated to the inner grid)                  src/CDMMergeUtils.cc:214-246

The three interpolations of one merged slice -- outer -> inner grid (for the border blend), blended inner -> target grid,
outer -> target grid -- run on the GPU through `Interpolator`; the blend and the overlay are the reference's own per-value
rules on the host (a few thousand values; they are not part of the gather path).  As in interpolator.py the CDM side (which
variables, coordinate systems, units) stays with the host application: grids are described by proj4 string and axes.
"""
from __future__ import annotations

import numpy as np

from .capi import FimexB200Error, Method
from .interpolator import Interpolator


def extend_inner_axis(inner, outer):
    """CDMMergerPrivate::extendInnerAxis (:233-274): the inner axis continued with its own step over the outer axis' range"""
    vi, vo = np.asarray(inner, dtype=np.float64), np.asarray(outer, dtype=np.float64)
    if vi.size < 2:
        raise FimexB200Error("no data for axis in inner")
    if vo.size < 2:
        raise FimexB200Error("no data for axis in outer")
    step_i, step_o = vi[1] - vi[0], vo[1] - vo[0]
    close = lambda a, b: abs(a - b) <= 1e-5 * max(abs(a), abs(b), 1e-30)  # `equal()` of the reference: relative comparison
    if not all(close(d, step_i) for d in np.diff(vi)):
        raise FimexB200Error("axis in inner does not have constant step size, cannot merge")
    if not all(close(d, step_o) for d in np.diff(vo)):
        raise FimexB200Error("axis in outer does not have constant step size, cannot merge")
    min_i, max_i = (vi[0], vi[-1]) if step_i > 0 else (vi[-1], vi[0])
    min_o, max_o = (vo[0], vo[-1]) if step_o > 0 else (vo[-1], vo[0])
    if min_i < min_o or max_i > max_o:
        raise FimexB200Error("top not inside  bottom")
    before = []
    n = vi[0] - step_i
    while min_o <= n <= max_o:
        before.append(n)
        n -= step_i
    after = []
    n = vi[-1] + step_i
    while min_o <= n <= max_o:
        after.append(n)
        n += step_i
    return np.concatenate([np.array(before[::-1], dtype=np.float64), vi, np.array(after, dtype=np.float64)])


def linear_border_smoothing(inner, outer, transition_width=5, border_width=2, use_outer_if_inner_undefined=True):
    """CDMBorderSmoothing::getDataSlice + CDMBorderSmoothing_Linear (defaults :55 of the header): `inner` and `outer` are
    [.., ny, nx] float64 arrays ON THE INNER GRID (NaN = undefined); outside the border the outer value, in the interior the inner
    one, a linear ramp over `transition_width` cells in between (corner distances are Euclidean)"""
    vi, vo = np.asarray(inner, dtype=np.float64), np.asarray(outer, dtype=np.float64)
    ny, nx = vi.shape[-2:]
    x = np.arange(nx)[None, :] + np.zeros((ny, 1), dtype=np.int64)
    y = np.arange(ny)[:, None] + np.zeros((1, nx), dtype=np.int64)
    xmin1, ymin1 = border_width, border_width
    xmax1, ymax1 = xmin1 + transition_width, ymin1 + transition_width
    xmax2, ymax2 = nx - border_width, ny - border_width
    xmin2, ymin2 = xmax2 - transition_width, ymax2 - transition_width
    dx = np.where(x < xmax1, xmax1 - x, np.where(x >= xmin2, x - xmin2, 0)).astype(np.float64)
    dy = np.where(y < ymax1, ymax1 - y, np.where(y >= ymin2, y - ymin2, 0)).astype(np.float64)
    alpha = np.clip(np.sqrt(dx * dx + dy * dy) / transition_width, 0.0, 1.0)  # one of dx, dy is 0 along the edges: dist == the other
    outside = (x < xmin1) | (x >= xmax2) | (y < ymin1) | (y >= ymax2)
    inside = (x >= xmax1) & (x < xmin2) & (y >= ymax1) & (y < ymin2)
    blended = vi + alpha * (vo - vi)
    blended = np.where(vo - vi == 0, vo, blended)
    merged = np.where(outside, vo, np.where(inside, vi, blended))
    merged = np.where(np.isnan(vo), vi, merged)  # outer undefined (or no smoothing): the inner value
    return np.where(np.isnan(vi), vo if use_outer_if_inner_undefined else np.nan, merged)


class Merger:
    """CDMMerger for one variable's horizontal coordinate systems: an inner (fine, small) grid merged into an outer (coarse,
    large) one on a target grid.

    proj_inner / proj_outer : proj4 strings; x/y axes in the projection's unit, degrees when `is_degree`
    """

    def __init__(self, proj_inner, x_inner, y_inner, proj_outer, x_outer, y_outer, is_degree):
        self.proj_inner, self.proj_outer = proj_inner, proj_outer
        self.x_inner, self.y_inner = np.asarray(x_inner, dtype=np.float64), np.asarray(y_inner, dtype=np.float64)
        self.x_outer, self.y_outer = np.asarray(x_outer, dtype=np.float64), np.asarray(y_outer, dtype=np.float64)
        self.is_degree = bool(is_degree)
        self.gridInterpolationMethod = Method.BILINEAR  # src/CDMMerger.cc:84
        self.transitionWidth, self.borderWidth = 5, 2   # CDMBorderSmoothing_LinearFactory defaults
        self.useOuterIfInnerUndefined = True
        self._outer_to_inner = None
        self._inner_to_target = None
        self._outer_to_target = None

    def setGridInterpolationMethod(self, method):
        self.gridInterpolationMethod = Method(int(method))

    def setTargetGrid(self, proj, tx, ty, tx_unit="m", ty_unit="m"):
        """CDMMerger::setTargetGrid -> makeCDM (:211-229)"""
        unit = "degree" if self.is_degree else "m"
        m = self.gridInterpolationMethod
        # readerSmooth: outer interpolated onto the INNER grid (makeMergedCDM, CDMMergeUtils.cc:234-243)
        self._outer_to_inner = Interpolator(self.proj_outer, self.x_outer, self.y_outer, self.is_degree)
        self._outer_to_inner.changeProjection(m, self.proj_inner, self.x_inner, self.y_inner, unit, unit)
        # interpolatedST: the smoothed inner field onto the target grid
        self._inner_to_target = Interpolator(self.proj_inner, self.x_inner, self.y_inner, self.is_degree)
        self._inner_to_target.changeProjection(m, proj, tx, ty, tx_unit, ty_unit)
        # readerOverlay: the outer field onto the target grid, underneath
        self._outer_to_target = Interpolator(self.proj_outer, self.x_outer, self.y_outer, self.is_degree)
        self._outer_to_target.changeProjection(m, proj, tx, ty, tx_unit, ty_unit)
        self.target_shape = (len(np.atleast_1d(ty)) if not isinstance(ty, str) else None, len(np.atleast_1d(tx)) if not isinstance(tx, str) else None)
        return self

    def setTargetGridFromInner(self):
        """CDMMerger::setTargetGridFromInner (:153-196): the inner grid extended over the outer grid's range, inner projection"""
        if self.proj_inner is None or self.proj_outer is None:
            raise FimexB200Error("extending grid failed, no inner variable with CS found")
        unit = "degree" if self.is_degree else "m"
        vx = extend_inner_axis(self.x_inner, self.x_outer)
        vy = extend_inner_axis(self.y_inner, self.y_outer)
        self.target_x, self.target_y = vx, vy
        return self.setTargetGrid(self.proj_inner, vx, vy, unit, unit)

    def getDataSlice(self, inner, outer):
        """One merged slice: `inner` [.., nyI, nxI] and `outer` [.., nyO, nxO] as float64 with NaN = undefined (getScaledDataSlice);
        returns [.., nyT, nxT] float64"""
        if self._outer_to_target is None:
            raise FimexB200Error("must call setTargetGrid or setTargetGridFromInner before getDataSlice")
        nan = float("nan")
        vi, vo = np.asarray(inner, dtype=np.float64), np.asarray(outer, dtype=np.float64)
        outer_on_inner = self._outer_to_inner.getDataSlice(vo, bad_value=nan)
        smooth = linear_border_smoothing(vi, outer_on_inner, self.transitionWidth, self.borderWidth, self.useOuterIfInnerUndefined)
        top = self._inner_to_target.getDataSlice(smooth, bad_value=nan)
        base = self._outer_to_target.getDataSlice(vo, bad_value=nan)
        return np.where(np.isnan(top), base, top)  # CDMOverlay.cc:80-85
