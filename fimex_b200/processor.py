"""Host-side mirror of the two CDMProcessor operations that share the vector-rotation kernels of the regridding path.

Reference: /root/reference/src/CDMProcessor.cc
  makeCachedVectorReprojection                              :99-145   (matrix of the grid's own axes)
  rotateVectorToLatLon(toLatLon, varNameX, varNameY, ...)    :367-418  (registers x/y pairs per coordinate system)
  rotateDirectionToLatLon(toLatLon, varNames)                :420-456
  getDataSlice, rotation branches                            :579-617 (vectors), :619-637 (directions)

SURVEY.md 8f rank 4 lists these as "free beneficiaries" of the path's tables and kernels.  As in interpolator.py the CDM
side (finding the coordinate system, the variable pairs and their fill values) stays with the host application; the grid is
described by its proj4 string and axes.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .cached import CachedVectorReprojection
from .capi import FimexB200Error


class Processor:
    """rotateVectorToLatLon / rotateDirectionToLatLon for one horizontal coordinate system.

    proj4        : Projection::getProj4String() of the grid
    x_axis,y_axis: the grid's axes in metres, or degrees when `is_degree` (lat/long and rotated lat/long)
    """

    def __init__(self, proj4, x_axis, y_axis, is_degree):
        self.proj4 = proj4
        self.x_axis = np.asarray(x_axis, dtype=np.float64)
        self.y_axis = np.asarray(y_axis, dtype=np.float64)
        self.is_degree = bool(is_degree)
        self.cachedVectorReprojection = None

    def rotateVectorToLatLon(self, toLatLon=True):
        """CDMProcessor::rotateVectorToLatLon (:367-418) -> makeCachedVectorReprojection (:99-145)"""
        self.cachedVectorReprojection = CachedVectorReprojection.fromGrid(capi.MIFI_VECTOR_KEEP_SIZE, self.proj4, self.x_axis, self.y_axis,
                                                                          self.is_degree, toLatLon)
        return self

    rotateDirectionToLatLon = rotateVectorToLatLon  # :420-456 builds the same matrix

    def getVectorSlices(self, x_data, y_data, x_bad_value=None, y_bad_value=None):
        """Both components of the rotation branch of getDataSlice (:579-617) in one pass; the reference rotates the pair once
        per requested component and keeps one half each time."""
        cvr = self._cvr()
        x_data, y_data = np.asarray(x_data), np.asarray(y_data)
        if x_data.size != y_data.size:
            raise FimexB200Error("xData != yData in vectorInterpolation")
        if x_bad_value is None:
            x_bad_value = capi.default_fill_value(x_data.dtype)
        if y_bad_value is None:
            y_bad_value = capi.default_fill_value(y_data.dtype)
        return cvr.getVectorSlice(x_data, y_data, x_bad_value, y_bad_value)

    def getDataSlice(self, data, counterpart, direction="x", bad_value=None, counterpart_bad_value=None):
        """getDataSlice(varName) for a variable registered as the x (or y) half of a pair: returns that half only"""
        if "x" in direction:
            return self.getVectorSlices(data, counterpart, bad_value, counterpart_bad_value)[0]
        if "y" in direction:
            return self.getVectorSlices(counterpart, data, counterpart_bad_value, bad_value)[1]
        raise FimexB200Error(f"could not find x,y direction for vector, direction: {direction}")

    def getDirectionSlice(self, angles):
        """The direction branch of getDataSlice (:619-637) for float angles in degrees without packing (scale 1, offset 0)"""
        cvr = self._cvr()
        out = np.array(angles, dtype=np.float32, copy=True)
        cvr.reprojectDirectionValues(out)
        return out

    def _cvr(self):
        if self.cachedVectorReprojection is None:
            raise FimexB200Error("no cached vector reprojection: call rotateVectorToLatLon first")
        return self.cachedVectorReprojection
