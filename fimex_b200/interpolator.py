"""Host-side mirror of the part of CDMInterpolator that sits on the hot path.

Reference: /root/reference/src/CDMInterpolator.cc
  changeProjection(method, proj_input, out_x_axis, out_y_axis, units)          :420-450   (dispatch on method)
  changeProjectionByProjectionParameters / ...ByCoordinates / ...ByForwardInterpolation  :1422-1503 / :1336-1420 / :1242-1334
  getDataSlice                                                                   :235-287   (per-slice driver)
  data2InterpolationArray / interpolationArray2Data                              :115-124   (bad <-> NaN adapters)
and the axis-string grammar of SpatialAxisSpec (src/SpatialAxisSpec.cc:84-150, include/fimex/Utils.h:373-407).

The CDM metadata side of the reference (coordinate-system discovery, CDM rewrite, NetCDF readers) stays with
the host application; here the source grid is described directly by its proj4 string and axes (or 2-D lon/lat),
which is exactly what the reference extracts from the CDM before it calls the functions this package replaces.
All array arithmetic happens in libfimex_b200.so on the GPU.
"""
from __future__ import annotations

import re

import numpy as np

from . import capi
from .cached import CachedForwardInterpolation, CachedInterpolation, CachedVectorReprojection
from .capi import FimexB200Error, Method

_DEGREE = re.compile(r".*degree.*", re.S)  # boost::regex degree(".*degree.*"), CDMInterpolator.cc:1443

MIFI_WGS84_LATLON_PROJ4 = "+proj=latlong +datum=WGS84 +towgs84=0,0,0 +no_defs"  # include/fimex/CDMconstants.h:118
MIFI_EARTH_RADIUS_M = 6371000  # include/fimex/CDMconstants.h:113


def tokenize_dotted(spec: str, delimiter: str = ",") -> list:
    """tokenizeDotted<double> (include/fimex/Utils.h:373-407): "0,0.5,...,3" -> [0, 0.5, 1, ..., 3]"""
    toks = [t for t in spec.split(delimiter)]
    vals: list = []
    i = 0
    while i < len(toks):
        cur = toks[i].strip()
        if cur == "...":
            if len(vals) < 2:
                raise FimexB200Error("tokenizeDotted: cannot use ... expansion, need at least two values before")
            last = vals[-1]
            dist = last - vals[-2]
            cur_val = last + dist
            direction = 1.0 if dist > 0 else -1.0
            i += 1
            if i < len(toks):
                after = float(toks[i])
                round_error = direction * dist * -1.0e-5
                while (cur_val - after) * direction < round_error:
                    vals.append(cur_val)
                    cur_val += dist
                vals.append(after)
        else:
            vals.append(float(cur))
        i += 1
    return vals


_RELATIVE_FINAL = re.compile(r"x\s*([+-])?\s*(\d\.?\d*)?\s*")  # TranslateRelativePlace, src/SpatialAxisSpec.cc:58


def axis_spec_requires_start_end(spec) -> bool:
    """SpatialAxisSpec::requireStartEnd (src/SpatialAxisSpec.cc:153-158) before setStartEnd"""
    return isinstance(spec, str) and "relativeStart" in spec


def spatial_axis_spec(spec, start=None, end=None) -> np.ndarray:
    """SpatialAxisSpec (src/SpatialAxisSpec.cc:84-150): absolute axis strings ("0,0.5,...,3"), or strings relative to the
    bounding box [start, end] of the data in the target projection ("0,50000,...,x,x+50000;relativeStart=0": 0 is the
    first multiple of the step inside the box, x the last one; test/testSpatialAxisSpec.cc)."""
    if not isinstance(spec, str):
        return np.asarray(spec, dtype=np.float64)
    steps = None
    relative = ""
    for part in spec.split(";"):
        kv = part.split("=")
        if len(kv) == 1:
            if steps is not None:
                raise FimexB200Error(f"axis-steps redefined from {steps} to {kv[0]}")
            steps = kv[0]
        elif len(kv) == 2 and kv[0] == "relativeStart":
            relative = kv[1]
        elif len(kv) == 2 and kv[0] == "unit":
            raise FimexB200Error("unit not supported yet in SpatialAxisSpec, please enter values in m or degree")
        elif len(kv) == 2:
            raise FimexB200Error(f"unknown axisSpec parameter: '{kv[0]}'")
        else:
            raise FimexB200Error(f"unknown axisSpec section: '{part}'")
    if relative == "":
        return np.asarray(tokenize_dotted(steps or ""), dtype=np.float64)
    if start is None or end is None:
        raise FimexB200Error("require start and end for axisSpec: " + spec)
    places = [p for p in (steps or "").split(",")]
    if len(places) < 2:
        raise FimexB200Error("SpatialAxisSpec requires at least 2 values with relative start definition, got: " + (steps or ""))
    delta = float(places[1]) - float(places[0])
    start_offset = int(start / delta) * delta  # static_cast<int>: towards zero
    if start < start_offset:
        start_offset += delta
    final = int(end / delta) * delta
    absolute = []
    for value in places:
        if value.strip() == "...":
            absolute.append("...")
            continue
        m = _RELATIVE_FINAL.search(value)
        if m:
            amount = float(m.group(2)) if m.group(2) else 0.0
            v = final + amount if m.group(1) == "+" else final - amount if m.group(1) == "-" else final
        else:
            v = start_offset + float(value)
        absolute.append(repr(float(v)))
    return np.asarray(tokenize_dotted(",".join(absolute)), dtype=np.float64)


def lon_lat_vals_to_matrix(lon_vals, lat_vals):
    """lonLatVals2Matrix (CDMInterpolator.cc:1226-1239): index ix + iy*lonSize"""
    lon_vals = np.asarray(lon_vals, dtype=np.float64)
    lat_vals = np.asarray(lat_vals, dtype=np.float64)
    lon2d = np.tile(lon_vals, lat_vals.size)
    lat2d = np.repeat(lat_vals, lon_vals.size)
    return lon2d, lat2d


class Interpolator:
    """The hot-path half of CDMInterpolator for one horizontal coordinate system.

    source_proj4 : proj4 string of the source CRS (Projection::getProj4String, ProjectionImpl.cc:116-137)
    x_axis,y_axis: source axes in the projection's unit ("m") or degrees when `is_degree`
                   (Projection::isDegree(): lat/long and rotated lat/long)
    lon2d, lat2d : optional 2-D coordinates in degrees (index ix + iy*nx) for coord_* and forward_* methods;
                   derived from the axes when the source is a lat/long grid (latLonProj, :1376-1380)
    has_xy_vectors: whether the data contains x/y-directed vector pairs (hasXYSpatialVectors(), :1490)
    """

    def __init__(self, source_proj4, x_axis, y_axis, is_degree, lon2d=None, lat2d=None, has_xy_vectors=False, x_dim="x", y_dim="y"):
        self.source_proj4 = source_proj4
        self.x_axis = np.asarray(x_axis, dtype=np.float64)
        self.y_axis = np.asarray(y_axis, dtype=np.float64)
        self.is_degree = bool(is_degree)
        self.lon2d = None if lon2d is None else np.asarray(lon2d, dtype=np.float64).ravel()
        self.lat2d = None if lat2d is None else np.asarray(lat2d, dtype=np.float64).ravel()
        self.has_xy_vectors = has_xy_vectors
        self.x_dim, self.y_dim = x_dim, y_dim
        self.cachedInterpolation = None
        self.cachedVectorReprojection = None
        self.method = Method.UNKNOWN
        self.distanceOfInterest = 0.0  # CDMInterpolator::setDistanceOfInterest (metres; <= 0: derived from the output axes)
        self._pair_cache = {}  # parked counterpart halves of x/y vector pairs, see getDataSlice(pair_key=...)

    # ---- setup --------------------------------------------------------------------------------
    def changeProjection(self, method, proj_input, out_x_axis, out_y_axis, out_x_axis_unit="m", out_y_axis_unit="m"):
        """CDMInterpolator::changeProjection (:345-450).  `method` may be the option string
        (--interpolate.method) or the enum value; axes may be option strings ("0,0.5,...,10") or arrays."""
        if isinstance(method, str):
            m = capi.mifi_string_to_interpolation_method(method)
        else:
            m = int(method)
        if axis_spec_requires_start_end(out_x_axis) or axis_spec_requires_start_end(out_y_axis):
            (x_min, x_max), (y_min, y_max) = self._bounding_box_in(proj_input)
            out_x = spatial_axis_spec(out_x_axis, x_min, x_max)
            out_y = spatial_axis_spec(out_y_axis, y_min, y_max)
        else:
            out_x = spatial_axis_spec(out_x_axis)
            out_y = spatial_axis_spec(out_y_axis)
        self.clearPairCache()
        self.cachedInterpolation = None
        self.cachedVectorReprojection = None
        if m in (Method.NEAREST_NEIGHBOR, Method.BILINEAR, Method.BICUBIC):
            self._by_projection_parameters(m, proj_input, out_x, out_y, out_x_axis_unit, out_y_axis_unit)
        elif m in (Method.COORD_NN, Method.COORD_NN_KD):
            self._by_coordinates(m, proj_input, out_x, out_y, out_x_axis_unit, out_y_axis_unit)
        elif Method.FORWARD_SUM <= m <= Method.FORWARD_UNDEF_MIN:
            self._by_forward_interpolation(m, proj_input, out_x, out_y, out_x_axis_unit, out_y_axis_unit)
        else:
            raise FimexB200Error(f"unknown projection method: {m}")
        self.method = Method(m)
        return self

    def _bounding_box_in(self, proj_input):
        """the bounding box of the data in the target projection for relative axis specs (CDMInterpolator.cc:349-401): all
        longitude/latitude values of the source grid projected to `proj_input`, min / max per axis"""
        m = re.search(r"\+proj=(\S+)", proj_input)
        name = m.group(1) if m else ""
        if name == "latlong":
            raise FimexB200Error("changeProjection with autotuning axes only implemented for projections in m, not degree yet")
        if self.lon2d is not None and self.lat2d is not None:
            lon, lat = self.lon2d, self.lat2d
        elif self.is_degree and "ob_tran" not in self.source_proj4:
            lon, lat = lon_lat_vals_to_matrix(self.x_axis, self.y_axis)  # 1-D axes made 2-D (:365-379)
        else:  # the reference reads the CDM's longitude/latitude variables; here they follow from the source grid itself
            lon, lat = self.convertToLonLat(np.tile(self.x_axis, self.y_axis.size), np.repeat(self.y_axis, self.x_axis.size))
        rc, x, y = capi.mifi_project_values(MIFI_WGS84_LATLON_PROJ4, proj_input, np.radians(lon), np.radians(lat))
        if rc != capi.MIFI_OK:
            raise FimexB200Error(f"unable to project axes from {MIFI_WGS84_LATLON_PROJ4} to {proj_input}")
        ok = np.isfinite(x) & np.isfinite(y)
        box = [(float(x[ok].min()), float(x[ok].max())), (float(y[ok].min()), float(y[ok].max()))]
        if name == "ob_tran":
            box = [(np.degrees(a), np.degrees(b)) for a, b in box]
        return box

    def _by_projection_parameters(self, method, proj_input, out_x, out_y, xunit, yunit):
        # :1440-1503
        x_deg = bool(_DEGREE.match(xunit))
        y_deg = bool(_DEGREE.match(yunit))
        ci = CachedInterpolation.fromProjection(method, proj_input, out_x, out_y, x_deg, y_deg, self.source_proj4, self.x_axis, self.y_axis,
                                                self.is_degree, self.x_dim, self.y_dim)
        ci.createReducedDomain(self.x_dim, self.y_dim)  # :1483-1485 (vector pairs share the horizontal id here)
        self.cachedInterpolation = ci
        if self.has_xy_vectors:  # :1490-1500; note the ORIGINAL (unconverted) axes and their types go in
            xt = capi.LONGITUDE if x_deg else capi.PROJ_AXIS
            yt = capi.LATITUDE if y_deg else capi.PROJ_AXIS
            self.cachedVectorReprojection = CachedVectorReprojection.fromProjection(capi.MIFI_VECTOR_KEEP_SIZE, self.source_proj4, proj_input,
                                                                                    out_x, out_y, xt, yt)

    def changeProjectionToLonLatValues(self, method, lon_vals, lat_vals):
        """CDMInterpolator::changeProjection(method, lonVals, latVals) (:460-510): a list of geographic points as the target
        (outX = n, outY = 1).  The reference narrows the values to float before use (double_to_float_cast, :454-457)."""
        lon_vals = np.asarray(lon_vals, dtype=np.float64).ravel()
        lat_vals = np.asarray(lat_vals, dtype=np.float64).ravel()
        if lon_vals.size != lat_vals.size:  # the reference logs an error and leaves the projection unchanged (:465-469)
            raise FimexB200Error(f"changeProjection, number of longitude and latitude values differs: {lon_vals.size} != {lat_vals.size}")
        lon32 = lon_vals.astype(np.float32).astype(np.float64)
        lat32 = lat_vals.astype(np.float32).astype(np.float64)
        return self.changeProjectionToTemplate(method, lon32, lat32, lon_vals.size, 1)

    # ---- Projection::convertFromLonLat / convertToLonLat (src/coordSys/Projection.cc:73-140) ------------------------------
    def _latlong_of_source(self):
        """"+proj=latlong " + getProj4EarthString() (ProjectionImpl.cc:139-159): a geographic CRS on the source's own earth
        figure.  The reference assembles it from the CF grid-mapping attributes; here it is read back from the proj4 string."""
        earth = [tok for tok in self.source_proj4.split()
                 if tok.split("=")[0] in ("+a", "+b", "+rf", "+e", "+es", "+f", "+R", "+ellps", "+datum", "+towgs84")]
        if not earth:
            earth = [f"+a={MIFI_EARTH_RADIUS_M}", "+e=0"]  # default (:152)
        return "+proj=latlong " + " ".join(earth)

    def convertFromLonLat(self, lon, lat):
        """degrees -> source projection coordinates (m, or degrees when the source isDegree)"""
        x, y = np.radians(np.asarray(lon, dtype=np.float64)), np.radians(np.asarray(lat, dtype=np.float64))
        ll = self._latlong_of_source()
        if ll != self.source_proj4:
            rc, x, y = capi.mifi_project_values(ll, self.source_proj4, x, y)
            if rc != capi.MIFI_OK:
                raise FimexB200Error(f"convertFromLonLat: unable to convert from '{ll}' to '{self.source_proj4}'")
        return (np.degrees(x), np.degrees(y)) if self.is_degree else (x, y)

    def convertToLonLat(self, x, y):
        x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
        if self.is_degree:
            x, y = np.radians(x), np.radians(y)
        ll = self._latlong_of_source()
        if ll != self.source_proj4:
            rc, x, y = capi.mifi_project_values(self.source_proj4, ll, x, y)
            if rc != capi.MIFI_OK:
                raise FimexB200Error(f"convertToLonLat: unable to convert from '{self.source_proj4}' to '{ll}'")
        return np.degrees(x), np.degrees(y)

    def changeProjectionToCrossSections(self, method, cross_sections):
        """CDMInterpolator::changeProjectionToCrossSections (:512-632).  `cross_sections`: [(name, [(lon, lat), ...]), ...] in
        degrees.  Every leg is sampled in the SOURCE projection with about one point per grid cell (:564-582), the points go
        back to lon/lat and become one point list (outY = 1).  Afterwards `vcross_names` and `vcross_bnds` (first and last
        index, inclusive, per cross-section: the reference's vcross_bnds variable, :611-626) describe the list."""
        if self.x_axis.size < 2 or self.y_axis.size < 2:
            raise FimexB200Error("x- or y-axis sizes < 2 elements, not possible to interpolate")
        dx = self.x_axis[1] - self.x_axis[0]
        dy = self.y_axis[1] - self.y_axis[0]
        if dx == 0 or dy == 0:
            raise FimexB200Error("cross-section calculation: dx or dy derived from first two elements == 0")
        lon_vals, lat_vals, starts, names = [], [], [], []
        for name, lon_lat in cross_sections:
            if len(lon_lat) == 0:
                continue
            names.append(name)
            starts.append(len(lon_vals))
            if len(lon_lat) == 1:
                lon_vals.append(float(lon_lat[0][0]))
                lat_vals.append(float(lon_lat[0][1]))
                continue
            for i in range(1, len(lon_lat)):
                px, py = self.convertFromLonLat([lon_lat[i - 1][0], lon_lat[i][0]], [lon_lat[i - 1][1], lon_lat[i][1]])
                xd, yd = px[1] - px[0], py[1] - py[0]
                num = int(np.floor(max(abs(xd / dx), abs(yd / dy))))  # grid points between the two coordinates
                xs, ys = [], []
                if i == 1:  # the first point only once: it is the last point of the previous leg otherwise
                    xs.append(px[0])
                    ys.append(py[0])
                for j in range(1, num):
                    xs.append(px[0] + j * xd / num)
                    ys.append(py[0] + j * yd / num)
                xs.append(px[1])
                ys.append(py[1])
                lo, la = self.convertToLonLat(xs, ys)
                lon_vals.extend(lo.tolist())
                lat_vals.extend(la.tolist())
        if not names:
            raise FimexB200Error("no cross-section with coordinates")
        self.vcross_names = names
        self.vcross_bnds = [(starts[i], (starts[i + 1] if i + 1 < len(starts) else len(lon_vals)) - 1) for i in range(len(starts))]
        return self.changeProjectionToLonLatValues(method, lon_vals, lat_vals)

    def changeProjectionToTemplate(self, method, tmpl_lon, tmpl_lat, out_x=None, out_y=None):
        """CDMInterpolator::changeProjection(method, tmplReader, tmplRefVarName) (:651-720) after the template's 2-D
        longitude/latitude (degrees, [y][x]) have been read by the host application ->
        changeProjectionByProjectionParametersToLatLonTemplate (:1706-1823)."""
        m = capi.mifi_string_to_interpolation_method(method) if isinstance(method, str) else int(method)
        if m in (Method.COORD_NN, Method.COORD_NN_KD) or Method.FORWARD_SUM <= m <= Method.FORWARD_UNDEF_MIN:
            raise FimexB200Error(f"projection method: {m}, not supported")  # :706-714
        if m not in (Method.NEAREST_NEIGHBOR, Method.BILINEAR, Method.BICUBIC):
            raise FimexB200Error(f"unknown projection method: {m}")
        tmpl_lon = np.asarray(tmpl_lon, dtype=np.float64)
        tmpl_lat = np.asarray(tmpl_lat, dtype=np.float64)
        if out_x is None or out_y is None:
            out_y, out_x = tmpl_lon.shape
        self.clearPairCache()
        self.cachedInterpolation = None
        self.cachedVectorReprojection = None
        ci = CachedInterpolation.fromTemplate(m, MIFI_WGS84_LATLON_PROJ4, tmpl_lon.ravel(), tmpl_lat.ravel(), out_x, out_y, self.source_proj4,
                                              self.x_axis, self.y_axis, self.is_degree, self.x_dim, self.y_dim)
        ci.createReducedDomain(self.x_dim, self.y_dim)  # :1797-1799
        self.cachedInterpolation = ci
        if self.has_xy_vectors:  # :1805-1820: matrix at the template points, source CRS -> geographic
            self.cachedVectorReprojection = CachedVectorReprojection.fromPoints(capi.MIFI_VECTOR_KEEP_SIZE, self.source_proj4,
                                                                                MIFI_WGS84_LATLON_PROJ4, 0 if self.is_degree else 1,
                                                                                tmpl_lon.ravel(), tmpl_lat.ravel())
        self.method = Method(m)
        return self

    def _source_lonlat(self):
        if self.lon2d is not None and self.lat2d is not None:
            return self.lon2d, self.lat2d
        if self.is_degree and "ob_tran" not in self.source_proj4:
            return lon_lat_vals_to_matrix(self.x_axis, self.y_axis)  # latLonProj branch, :1376-1380
        raise FimexB200Error("coordinate-based interpolation needs 2-D longitude/latitude of the source grid")

    def _by_coordinates(self, method, proj_input, out_x, out_y, xunit, yunit):
        # :1336-1420
        lon2d, lat2d = self._source_lonlat()
        self.cachedInterpolation = CachedInterpolation.fromCoordinates(method, proj_input, out_x, out_y, bool(_DEGREE.match(xunit)),
                                                                       bool(_DEGREE.match(yunit)), lon2d, lat2d, self.x_axis.size,
                                                                       self.y_axis.size, self.x_dim, self.y_dim,
                                                                       maxDistance=self.distanceOfInterest)
        # no reduced domain, no vector rotation on this path (:1417-1419)

    def _by_forward_interpolation(self, method, proj_input, out_x, out_y, xunit, yunit):
        # :1242-1334
        lon2d, lat2d = self._source_lonlat()
        self.cachedInterpolation = CachedForwardInterpolation.fromCoordinates(method, proj_input, out_x, out_y, bool(_DEGREE.match(xunit)),
                                                                              bool(_DEGREE.match(yunit)), lon2d, lat2d, self.x_axis.size,
                                                                              self.y_axis.size, self.x_dim, self.y_dim)

    # ---- per slice ------------------------------------------------------------------------------
    def getDataSlice(self, data, bad_value=None, counterpart=None, direction="x", counterpart_bad_value=None, pair_key=None):
        """CDMInterpolator::getDataSlice for an in-memory slice [.., y, x] of the FULL source grid (:235-287):
        crop to the reduced domain, fill -> NaN, interpolate, [rotate with the counterpart component], NaN -> fill
        and cast back to the variable's type -- one C-ABI call, the adapters run inside the gather kernel.
        `bad_value` = CDM::getFillValue(varName): the _FillValue attribute, else the type's default.

        `pair_key` (any hashable naming the x/y pair and the slice, e.g. ("x_wind", "y_wind", unLimDimPos)): the reference
        interpolates and rotates BOTH components on each component's call (:259-276) and throws one away; with a key the
        other half is parked and handed out when the counterpart is asked for with the same key, so
        the pair costs one pass instead of two.  A parked half is given out once; `clearPairCache()` drops what is left."""
        ci = self.cachedInterpolation
        if ci is None:
            raise FimexB200Error("no cached interpolation: call changeProjection first")
        on_device = type(data).__module__.startswith("torch")  # a CUDA tensor: the slab stays on the GPU, result is a CUDA tensor
        if not on_device:
            data = np.asarray(data)
        if bad_value is None:
            bad_value = capi.default_fill_value(data.dtype)
        lead = tuple(data.shape[:-2])
        if counterpart is not None and self.cachedVectorReprojection is not None:
            if "x" in direction:
                want = 0
            elif "y" in direction:
                want = 1
            else:
                raise FimexB200Error(f"could not find x,y direction for vector, direction: {direction}")
            cache = self._pair_cache
            if pair_key is not None and (pair_key, want) in cache:
                return cache.pop((pair_key, want))
            arr = ci.getInputDataSlice(data)
            other = ci.getInputDataSlice(counterpart if on_device else np.asarray(counterpart, dtype=data.dtype))
            cbad = bad_value if counterpart_bad_value is None else counterpart_bad_value
            if want == 0:
                both = ci.getVectorSlice(arr, other, bad_value, cbad, self.cachedVectorReprojection)
            else:
                both = ci.getVectorSlice(other, arr, cbad, bad_value, self.cachedVectorReprojection)
            shape = lead + (ci.getOutY(), ci.getOutX())
            if pair_key is not None:
                while len(cache) >= self.pairCacheSlots:  # bounded: oldest first
                    cache.pop(next(iter(cache)))
                cache[(pair_key, 1 - want)] = both[1 - want].reshape(shape)
            return both[want].reshape(shape)
        arr = ci.getInputDataSlice(data)
        out = ci.getDataSlice(arr, bad_value)
        return out.reshape(lead + (ci.getOutY(), ci.getOutX()))

    pairCacheSlots = 4  # parked counterpart slices kept at most (each is one output slice)

    def clearPairCache(self):
        self._pair_cache.clear()
