"""Host-side mirrors of the reference's cached regridding operators, forwarding to the C ABI.

Reference classes (arebru/fimex 0.67.2):
  CachedInterpolationInterface / CachedInterpolation   include/fimex/CachedInterpolation.h:60-161,
                                                        src/CachedInterpolation.cc:93-200
  CachedForwardInterpolation                            src/CachedForwardInterpolation.h:34-61, .cc:62-131
  CachedVectorReprojection                              include/fimex/CachedVectorReprojection.h:33-63,
                                                        src/CachedVectorReprojection.cc:35-55

Same constructor arguments and method names.  ``interpolateValues`` takes/returns numpy arrays (host
buffers: the copies are inside the call, like a Fimex host would use it) or torch CUDA tensors
(device-resident slabs: no copy, no synchronisation).  Errors raise FimexB200Error (the reference throws
CDMException).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import FimexB200Error, Method, check, f32, f64, load, ptr


def _is_torch(a) -> bool:
    return not isinstance(a, np.ndarray) and hasattr(a, "data_ptr")


def _stream_ptr(stream):
    if stream is None:
        try:
            import torch
            return C.c_void_p(torch.cuda.current_stream().cuda_stream)
        except Exception:  # pragma: no cover
            return None
    return C.c_void_p(stream if isinstance(stream, int) else stream.cuda_stream)


class CachedInterpolationInterface:
    """include/fimex/CachedInterpolation.h:60-99"""

    def __init__(self, xDimName: str, yDimName: str):
        self._xDimName = xDimName
        self._yDimName = yDimName
        self._h = C.c_void_p()
        self._reduced = None  # (xDim, yDim, xMin, yMin, xOrg, yOrg) == ReducedInterpolationDomain

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            load().fb200_interp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- getters ----------------------------------------------------------------------------------
    def getInX(self) -> int:
        return int(load().fb200_interp_in_x(self._h))

    def getInY(self) -> int:
        return int(load().fb200_interp_in_y(self._h))

    def getOutX(self) -> int:
        return int(load().fb200_interp_out_x(self._h))

    def getOutY(self) -> int:
        return int(load().fb200_interp_out_y(self._h))

    def reducedDomain(self):
        return self._reduced

    def points(self):
        """(pointsOnXAxis, pointsOnYAxis) as held by the handle (after any crop), host copies"""
        n = self._npoints()
        px, py = np.empty(n), np.empty(n)
        check(load().fb200_interp_get_points(self._h, ptr(px), ptr(py)), "get_points")
        return px, py

    def device_points(self):
        """torch views (no copy) of the fp64 position tables on the device, for the NCCL broadcast"""
        import torch
        a, b, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        check(load().fb200_interp_device_points(self._h, C.byref(a), C.byref(b), C.byref(n)), "device_points")
        dev = load().fb200_get_device()

        def view(p):
            iface = {"shape": (n.value,), "typestr": "<f8", "data": (p.value, False), "version": 2}
            holder = type("_Arr", (), {"__cuda_array_interface__": iface})()
            return torch.as_tensor(holder, device=f"cuda:{dev}")

        return view(a), view(b)

    def _npoints(self):
        raise NotImplementedError

    def getInputDataSlice(self, array):
        """The array-level part of getInputDataSlice (CachedInterpolation.cc:44-90): crop [.., y, x] to the
        reduced domain if there is one (numpy array, or a CUDA tensor cropped on the device)."""
        if self._reduced is None:
            return array
        _, _, x0, y0, _, _ = self._reduced
        if _is_torch(array):
            return array[..., y0:y0 + self.getInY(), x0:x0 + self.getInX()].contiguous()
        return np.ascontiguousarray(array[..., y0:y0 + self.getInY(), x0:x0 + self.getInX()])

    # -- the hot path -----------------------------------------------------------------------------
    def interpolateValues(self, inData, out=None, stream=None):
        """interpolateValues(inData, size, newSize): [inZ][inY][inX] -> new array [inZ][outY][outX]."""
        lib = load()
        new_size = C.c_size_t()
        if _is_torch(inData):
            import torch
            if not inData.is_cuda or inData.dtype != torch.float32 or not inData.is_contiguous():
                raise FimexB200Error("device path needs a contiguous float32 CUDA tensor")
            size = inData.numel()
            n = int(lib.fb200_interp_new_size(self._h, size))
            if out is None:
                out = torch.empty(n, dtype=torch.float32, device=inData.device)
            elif out.numel() < n or not out.is_contiguous():
                raise FimexB200Error("output tensor too small or not contiguous")
            check(lib.fb200_interp_interpolate_values_device(self._h, ptr(inData), size, ptr(out), C.byref(new_size), _stream_ptr(stream)),
                  "interpolateValues")
            nz = n // max(1, self.getOutX() * self.getOutY())
            return out[:n].view(nz, self.getOutY(), self.getOutX())
        a = f32(inData)
        n = int(lib.fb200_interp_new_size(self._h, a.size))
        if out is None:
            out = np.empty(n, dtype=np.float32)
        elif out.size < n or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise FimexB200Error("output array too small, not float32 or not contiguous")
        check(lib.fb200_interp_interpolate_values(self._h, ptr(a), a.size, ptr(out), C.byref(new_size)), "interpolateValues")
        assert new_size.value == n
        nz = n // max(1, self.getOutX() * self.getOutY())
        return out.reshape(-1)[:n].reshape(nz, self.getOutY(), self.getOutX())


    def addPreprocess(self, procString: str):
        """CDMInterpolator::addPreprocess with the --interpolate.preprocess option string, e.g. "fill2d(0.01,1.6,100)" """
        check(load().fb200_interp_add_preprocess(self._h, procString.encode()), "addPreprocess")

    def addPostprocess(self, procString: str):
        check(load().fb200_interp_add_postprocess(self._h, procString.encode()), "addPostprocess")

    # -- the whole slice body of CDMInterpolator::getDataSlice in one call ---------------------------
    def getDataSlice(self, inData, badValue, outType=None, stream=None, out=None):
        """data2InterpolationArray -> interpolateValues -> interpolationArray2Data (CDMInterpolator.cc:250-258, 284-285):
        `inData` [inZ][inY][inX] of any CDM numeric type (numpy array or CUDA tensor), `badValue` = CDM::getFillValue of the
        variable, result [inZ][outY][outX] in `outType` (default: the input's type)."""
        lib = load()
        new_size = C.c_size_t()
        if _is_torch(inData):
            import torch
            if not inData.is_cuda or not inData.is_contiguous():
                raise FimexB200Error("device path needs a contiguous CUDA tensor")
            out_dtype = outType if outType is not None else inData.dtype
            n = int(lib.fb200_interp_new_size(self._h, inData.numel()))
            if out is None:
                out = torch.empty(n, dtype=out_dtype, device=inData.device)
            elif out.numel() < n or out.dtype != out_dtype or not out.is_contiguous():
                raise FimexB200Error("output tensor too small, of the wrong type or not contiguous")
            out = out.view(-1)[:n]
            check(lib.fb200_interp_get_data_slice_device(self._h, int(capi.cdm_type(inData.dtype)), ptr(inData), inData.numel(), float(badValue),
                                                         int(capi.cdm_type(out_dtype)), ptr(out), C.byref(new_size), _stream_ptr(stream)),
                  "getDataSlice")
        else:
            a = np.ascontiguousarray(inData)
            out_dtype = np.dtype(outType) if outType is not None else a.dtype
            n = int(lib.fb200_interp_new_size(self._h, a.size))
            if out is None:
                out = np.empty(n, dtype=out_dtype)
            elif out.size < n or out.dtype != out_dtype or not out.flags.c_contiguous:
                raise FimexB200Error("output array too small, of the wrong type or not contiguous")
            out = out.reshape(-1)[:n]
            check(lib.fb200_interp_get_data_slice(self._h, int(capi.cdm_type(a.dtype)), ptr(a), a.size, float(badValue),
                                                  int(capi.cdm_type(out_dtype)), ptr(out), C.byref(new_size)), "getDataSlice")
        nz = n // max(1, self.getOutX() * self.getOutY())
        return out.reshape(nz, self.getOutY(), self.getOutX())


class CachedInterpolation(CachedInterpolationInterface):
    """CachedInterpolation(xDimName, yDimName, funcType, pointsOnXAxis, pointsOnYAxis, inX, inY, outX, outY)
    -- include/fimex/CachedInterpolation.h:126-128"""

    def __init__(self, xDimName, yDimName, funcType, pointsOnXAxis, pointsOnYAxis, inX, inY, outX, outY):
        super().__init__(xDimName, yDimName)
        lib = load()
        if _is_torch(pointsOnXAxis):
            check(lib.fb200_cached_interpolation_create_device(int(funcType), ptr(pointsOnXAxis), ptr(pointsOnYAxis), inX, inY, outX, outY,
                                                               C.byref(self._h)), "CachedInterpolation")
        else:
            px, py = f64(pointsOnXAxis), f64(pointsOnYAxis)
            if px.size != outX * outY or py.size != outX * outY:
                raise FimexB200Error("pointsOnXAxis/pointsOnYAxis must hold outX*outY values")
            check(lib.fb200_cached_interpolation_create(int(funcType), ptr(px), ptr(py), inX, inY, outX, outY, C.byref(self._h)),
                  "CachedInterpolation")

    @classmethod
    def _wrap(cls, handle, xDimName="x", yDimName="y"):
        self = cls.__new__(cls)
        CachedInterpolationInterface.__init__(self, xDimName, yDimName)
        self._h = handle
        return self

    @classmethod
    def fromProjection(cls, funcType, proj_target, out_x_axis, out_y_axis, out_x_is_degree, out_y_is_degree, proj_source, in_x_axis,
                       in_y_axis, in_is_degree, xDimName="x", yDimName="y"):
        """index tables of changeProjectionByProjectionParameters (CDMInterpolator.cc:1440-1481), built on the device"""
        ox, oy, ix, iy = f64(out_x_axis), f64(out_y_axis), f64(in_x_axis), f64(in_y_axis)
        h = C.c_void_p()
        check(load().fb200_cached_interpolation_create_from_projection(int(funcType), proj_target.encode(), ptr(ox), ptr(oy), ox.size, oy.size,
                                                                       int(out_x_is_degree), int(out_y_is_degree), proj_source.encode(), ptr(ix),
                                                                       ptr(iy), ix.size, iy.size, int(in_is_degree), C.byref(h)),
              "changeProjectionByProjectionParameters")
        return cls._wrap(h, xDimName, yDimName)

    @classmethod
    def fromTemplate(cls, funcType, proj_template, tmplLon, tmplLat, outX, outY, proj_source, in_x_axis, in_y_axis, in_is_degree,
                     xDimName="x", yDimName="y"):
        """index tables of changeProjectionByProjectionParametersToLatLonTemplate (CDMInterpolator.cc:1755-1803): the target is a
        2-D lon/lat template in degrees (or a point list with outY = 1)"""
        lo, la, ix, iy = f64(tmplLon), f64(tmplLat), f64(in_x_axis), f64(in_y_axis)
        if lo.size != outX * outY or la.size != outX * outY:
            raise FimexB200Error("template lon/lat must hold outX*outY values")
        h = C.c_void_p()
        check(load().fb200_cached_interpolation_create_from_template(int(funcType), proj_template.encode(), ptr(lo), ptr(la), outX, outY,
                                                                     proj_source.encode(), ptr(ix), ptr(iy), ix.size, iy.size,
                                                                     int(in_is_degree), C.byref(h)), "changeProjection (lat/lon template)")
        return cls._wrap(h, xDimName, yDimName)

    @classmethod
    def fromCoordinates(cls, funcType, proj_target, out_x_axis, out_y_axis, out_x_is_degree, out_y_is_degree, lon2d, lat2d, inX, inY,
                        xDimName="x", yDimName="y", maxDistance=0.0):
        """index tables of changeProjectionByCoordinates (CDMInterpolator.cc:1387-1412): coord_nearestneighbor, or coord_kdtree
        with `maxDistance` metres (the reference's setDistanceOfInterest; 0 = derived from the output axes)"""
        ox, oy, lo, la = f64(out_x_axis), f64(out_y_axis), f64(lon2d), f64(lat2d)
        if lo.size != inX * inY or la.size != inX * inY:
            raise FimexB200Error("lon2d/lat2d must hold inX*inY values")
        h = C.c_void_p()
        check(load().fb200_cached_interpolation_create_from_coordinates_kd(int(funcType), proj_target.encode(), ptr(ox), ptr(oy), ox.size,
                                                                           oy.size, int(out_x_is_degree), int(out_y_is_degree), ptr(lo),
                                                                           ptr(la), inX, inY, float(maxDistance), C.byref(h)),
              "changeProjectionByCoordinates")
        return cls._wrap(h, xDimName, yDimName)

    def _npoints(self):
        return self.getOutX() * self.getOutY()

    def createReducedDomain(self, xDimName=None, yDimName=None):
        """src/CachedInterpolation.cc:159-200"""
        red, x0, y0 = C.c_int(), C.c_longlong(), C.c_longlong()
        org = (self.getInX(), self.getInY())
        check(load().fb200_interp_create_reduced_domain(self._h, C.byref(red), C.byref(x0), C.byref(y0)), "createReducedDomain")
        if red.value and self._reduced is None:
            self._reduced = (xDimName or self._xDimName, yDimName or self._yDimName, x0.value, y0.value, org[0], org[1])
        return bool(red.value)

    def interpolateVector(self, uIn, vIn, vectorReprojection=None, stream=None):
        """Both components of an x/y vector in one pass, rotated in the same kernel
        (the fused form of CDMInterpolator.cc:255-276)."""
        lib = load()
        new_size = C.c_size_t()
        vh = vectorReprojection._h if vectorReprojection is not None else None
        if _is_torch(uIn):
            import torch
            size = uIn.numel()
            n = int(lib.fb200_interp_new_size(self._h, size))
            uo = torch.empty(n, dtype=torch.float32, device=uIn.device)
            vo = torch.empty(n, dtype=torch.float32, device=uIn.device)
            check(lib.fb200_interp_interpolate_vector_device(self._h, vh, ptr(uIn), ptr(vIn), size, ptr(uo), ptr(vo), C.byref(new_size),
                                                             _stream_ptr(stream)), "interpolateVector")
        else:
            u, v = f32(uIn), f32(vIn)
            if u.size != v.size:
                raise FimexB200Error("u and v differ in size")
            n = int(lib.fb200_interp_new_size(self._h, u.size))
            uo = np.empty(n, dtype=np.float32)
            vo = np.empty(n, dtype=np.float32)
            check(lib.fb200_interp_interpolate_vector(self._h, vh, ptr(u), ptr(v), u.size, ptr(uo), ptr(vo), C.byref(new_size)),
                  "interpolateVector")
        nz = n // max(1, self.getOutX() * self.getOutY())
        shape = (nz, self.getOutY(), self.getOutX())
        return uo.reshape(shape), vo.reshape(shape)


    def getVectorSlice(self, uIn, vIn, badValueU, badValueV, vectorReprojection=None, outType=None, stream=None):
        """The vector branch of CDMInterpolator::getDataSlice (:250-285) for both components at once: fill -> NaN,
        interpolate, rotate, NaN -> fill + cast."""
        lib = load()
        new_size = C.c_size_t()
        vh = vectorReprojection._h if vectorReprojection is not None else None
        if _is_torch(uIn):
            import torch
            if uIn.dtype != vIn.dtype or uIn.numel() != vIn.numel():
                raise FimexB200Error("u and v differ in type or size")
            out_dtype = outType if outType is not None else uIn.dtype
            n = int(lib.fb200_interp_new_size(self._h, uIn.numel()))
            uo = torch.empty(n, dtype=out_dtype, device=uIn.device)
            vo = torch.empty(n, dtype=out_dtype, device=uIn.device)
            check(lib.fb200_interp_get_vector_slice_device(self._h, vh, int(capi.cdm_type(uIn.dtype)), ptr(uIn), ptr(vIn), uIn.numel(),
                                                           float(badValueU), float(badValueV), int(capi.cdm_type(out_dtype)), ptr(uo), ptr(vo),
                                                           C.byref(new_size), _stream_ptr(stream)), "getVectorSlice")
        else:
            u, v = np.ascontiguousarray(uIn), np.ascontiguousarray(vIn)
            if u.dtype != v.dtype or u.size != v.size:
                raise FimexB200Error("u and v differ in type or size")
            out_dtype = np.dtype(outType) if outType is not None else u.dtype
            n = int(lib.fb200_interp_new_size(self._h, u.size))
            uo = np.empty(n, dtype=out_dtype)
            vo = np.empty(n, dtype=out_dtype)
            check(lib.fb200_interp_get_vector_slice(self._h, vh, int(capi.cdm_type(u.dtype)), ptr(u), ptr(v), u.size, float(badValueU),
                                                    float(badValueV), int(capi.cdm_type(out_dtype)), ptr(uo), ptr(vo), C.byref(new_size)),
                  "getVectorSlice")
        nz = n // max(1, self.getOutX() * self.getOutY())
        shape = (nz, self.getOutY(), self.getOutX())
        return uo.reshape(shape), vo.reshape(shape)


class CachedForwardInterpolation(CachedInterpolationInterface):
    """CachedForwardInterpolation(xDimName, yDimName, funcType, pOnX, pOnY, inX, inY, outX, outY)
    -- src/CachedForwardInterpolation.h:51-53"""

    def __init__(self, xDimName, yDimName, funcType, pOnX, pOnY, inX, inY, outX, outY):
        super().__init__(xDimName, yDimName)
        px, py = f64(pOnX), f64(pOnY)
        if px.size != inX * inY or py.size != inX * inY:
            raise FimexB200Error("pOnX/pOnY must hold inX*inY values")
        check(load().fb200_cached_forward_interpolation_create(int(funcType), ptr(px), ptr(py), inX, inY, outX, outY, C.byref(self._h)),
              "CachedForwardInterpolation")

    @classmethod
    def fromCoordinates(cls, funcType, proj_target, out_x_axis, out_y_axis, out_x_is_degree, out_y_is_degree, lon2d, lat2d, inX, inY,
                        xDimName="x", yDimName="y"):
        """index tables of changeProjectionByForwardInterpolation (CDMInterpolator.cc:1289-1332)"""
        ox, oy, lo, la = f64(out_x_axis), f64(out_y_axis), f64(lon2d), f64(lat2d)
        h = C.c_void_p()
        check(load().fb200_cached_forward_interpolation_create_from_coordinates(int(funcType), proj_target.encode(), ptr(ox), ptr(oy), ox.size,
                                                                                oy.size, int(out_x_is_degree), int(out_y_is_degree), ptr(lo),
                                                                                ptr(la), inX, inY, C.byref(h)),
              "changeProjectionByForwardInterpolation")
        self = cls.__new__(cls)
        CachedInterpolationInterface.__init__(self, xDimName, yDimName)
        self._h = h
        return self

    def _npoints(self):
        return self.getInX() * self.getInY()


class CachedVectorReprojection:
    """CachedVectorReprojection(method, matrix, ox, oy) -- include/fimex/CachedVectorReprojection.h:36-44"""

    def __init__(self, method=capi.MIFI_VECTOR_KEEP_SIZE, matrix=None, ox=0, oy=0):
        self._h = C.c_void_p()
        m = None if matrix is None else f64(matrix)
        if m is not None and m.size != 4 * ox * oy:
            raise FimexB200Error("matrix must hold 4*ox*oy values")
        check(load().fb200_vector_create(int(method), ptr(m), int(ox), int(oy), C.byref(self._h)), "CachedVectorReprojection")
        self.ox, self.oy = int(ox), int(oy)

    @classmethod
    def fromProjection(cls, method, proj_input, proj_output, out_x_axis, out_y_axis, out_x_axis_type, out_y_axis_type):
        """mifi_get_vector_reproject_matrix on the device (as called at CDMInterpolator.cc:1492-1500)"""
        ox, oy = f64(out_x_axis), f64(out_y_axis)
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        check(load().fb200_vector_create_from_projection(int(method), proj_input.encode(), proj_output.encode(), ptr(ox), ptr(oy),
                                                         int(out_x_axis_type), int(out_y_axis_type), ox.size, oy.size, C.byref(self._h)),
              "mifi_get_vector_reproject_matrix")
        self.ox, self.oy = ox.size, oy.size
        return self

    @classmethod
    def fromPoints(cls, method, proj_input, proj_output, inputIsMetric, lon, lat):
        """mifi_get_vector_reproject_matrix_points for a point list in degrees (lat/lon-template path, CDMInterpolator.cc:1805-1823)"""
        lo, la = f64(lon), f64(lat)
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        check(load().fb200_vector_create_from_points(int(method), proj_input.encode(), proj_output.encode(), int(inputIsMetric), ptr(lo),
                                                     ptr(la), lo.size, C.byref(self._h)), "mifi_get_vector_reproject_matrix_points")
        self.ox, self.oy = lo.size, 1
        return self

    @classmethod
    def fromGrid(cls, method, proj, x_axis, y_axis, is_degree, toLatLon=True):
        """makeCachedVectorReprojection(dataReader, cs, toLatLon) (src/CDMProcessor.cc:99-145): the rotation matrix of a grid's
        own axes towards geographic directions (toLatLon) or from geographic to grid directions"""
        xa, ya = f64(x_axis), f64(y_axis)
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        check(load().fb200_vector_create_from_grid(int(method), proj.encode(), ptr(xa), ptr(ya), xa.size, ya.size, int(bool(is_degree)),
                                                   int(bool(toLatLon)), C.byref(self._h)), "makeCachedVectorReprojection")
        self.ox, self.oy = xa.size, ya.size
        return self

    def getVectorSlice(self, uIn, vIn, badValueU, badValueV, outType=None, stream=None):
        """The rotation branch of CDMProcessor::getDataSlice (src/CDMProcessor.cc:579-617) for both components at once:
        fill -> NaN, reprojectValues, NaN -> fill + cast back to the variable's type"""
        lib = load()
        if _is_torch(uIn):
            import torch
            if uIn.dtype != vIn.dtype or uIn.numel() != vIn.numel():
                raise FimexB200Error("xData != yData in vectorInterpolation")
            out_dtype = outType if outType is not None else uIn.dtype
            uo = torch.empty(uIn.shape, dtype=out_dtype, device=uIn.device)
            vo = torch.empty(uIn.shape, dtype=out_dtype, device=uIn.device)
            check(lib.fb200_vector_get_slice_device(self._h, int(capi.cdm_type(uIn.dtype)), ptr(uIn), ptr(vIn), uIn.numel(), float(badValueU),
                                                    float(badValueV), int(capi.cdm_type(out_dtype)), ptr(uo), ptr(vo), _stream_ptr(stream)),
                  "rotateVectorToLatLon")
            return uo, vo
        u, v = np.ascontiguousarray(uIn), np.ascontiguousarray(vIn)
        if u.dtype != v.dtype or u.size != v.size:
            raise FimexB200Error("xData != yData in vectorInterpolation")  # CDMProcessor.cc:608-610
        out_dtype = np.dtype(outType) if outType is not None else u.dtype
        uo = np.empty(u.shape, dtype=out_dtype)
        vo = np.empty(u.shape, dtype=out_dtype)
        check(lib.fb200_vector_get_slice(self._h, int(capi.cdm_type(u.dtype)), ptr(u), ptr(v), u.size, float(badValueU), float(badValueV),
                                         int(capi.cdm_type(out_dtype)), ptr(uo), ptr(vo)), "rotateVectorToLatLon")
        return uo, vo

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            load().fb200_vector_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def getMatrix(self):
        m = np.empty(4 * self.ox * self.oy)
        check(load().fb200_vector_get_matrix(self._h, ptr(m)), "getMatrix")
        return m

    def reprojectValues(self, uValues, vValues, stream=None):
        """rotate in place (numpy float32 arrays or torch CUDA tensors) -- src/CachedVectorReprojection.cc:35-44"""
        lib = load()
        if _is_torch(uValues):
            check(lib.fb200_vector_reproject_values_device(self._h, ptr(uValues), ptr(vValues), uValues.numel(), _stream_ptr(stream)),
                  "reprojectValues")
            return
        for a in (uValues, vValues):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags.c_contiguous):
                raise FimexB200Error("reprojectValues works in place on contiguous float32 arrays")
        check(lib.fb200_vector_reproject_values(self._h, ptr(uValues), ptr(vValues), uValues.size), "reprojectValues")

    def reprojectDirectionValues(self, angles):
        if not (isinstance(angles, np.ndarray) and angles.dtype == np.float32 and angles.flags.c_contiguous):
            raise FimexB200Error("reprojectDirectionValues works in place on a contiguous float32 array")
        check(load().fb200_vector_reproject_direction_values(self._h, ptr(angles), angles.size), "reprojectDirectionValues")
