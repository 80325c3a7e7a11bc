#!/usr/bin/env python
"""Extract the inputs of the reference's CDMInterpolator / CDMProcessor tests (test/testInterpolator.cc, test/testProcessor.cc)
into one small fixture.

The GPU box has no /root/reference and no NetCDF reader, so the arrays the reference's tests read from its own classic
NetCDF-3 test files are stored in tests/golden/interpolator_fixtures.npz.  Only data is taken (coordinates, the fields the
tests look at, the grid mapping's proj4 attribute); run here once:

    python tests/golden/make_interpolator_fixtures.py [/root/reference]
"""
import os
import sys

import numpy as np
from scipy.io import netcdf_file


def native(a):
    a = np.asarray(a)
    return np.ascontiguousarray(a.astype(a.dtype.newbyteorder("=")))


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = {}

    # test_interpolator2coords (:129-180): temp2 lives on (y_c, x_c) with its own 2-D coordinates
    nc = netcdf_file(os.path.join(ref, "test/twoCoordsTest.nc"), "r", mmap=False)
    out["two_proj4"] = np.array(nc.variables["projection_1"].proj4.decode())
    for name in ("x_c", "y_c", "longitude2", "latitude2", "temp2"):
        out["two_" + name] = native(nc.variables[name].data)

    # test_interpolatorSatellite (:105-126): swath with 2-D lat/lon (float, _FillValue -999), no projection
    nc = netcdf_file(os.path.join(ref, "test/satellite_cma.nc"), "r", mmap=False)
    for name in ("lat", "lon", "cma"):
        out["sat_" + name] = native(nc.variables[name].data)

    # test_interpolator_template (:220-239) and test_interpolator_latlon (:241-264)
    nc = netcdf_file(os.path.join(ref, "test/erai.sfc.40N.0.75d.200301011200.nc"), "r", mmap=False)
    out["erai_proj4"] = np.array(nc.variables["projection_regular_ll"].proj4.decode())
    for name in ("longitude", "latitude", "ga_skt"):
        out["erai_" + name] = native(nc.variables[name].data)
    nc = netcdf_file(os.path.join(ref, "test/template_noaa17.nc"), "r", mmap=False)
    for name in ("longitude", "latitude"):
        out["tmpl_" + name] = native(nc.variables[name].data)

    # test_rotate (test/testProcessor.cc:71-93): rotateAllVectorsToLatLon on the 10 m wind of coordTest.nc, first time step
    nc = netcdf_file(os.path.join(ref, "test/coordTest.nc"), "r", mmap=False)
    out["coord_proj4"] = np.array(nc.variables["projection_1"].proj4.decode())
    for name in ("x", "y"):
        out["coord_" + name] = native(nc.variables[name].data)
    for name in ("x_wind_10m", "y_wind_10m"):
        out["coord_" + name] = native(nc.variables[name].data[0])

    # BASELINE config 1: the fimex CLI's bilinear regrid of test/hirlam12.nc to a 0.5-degree lat/long grid.  The file's own
    # data (axes in degrees, no grid_mapping => latitude_longitude on the default sphere, CF1_xCoordSysBuilder.cc:398-410),
    # plus golden outputs from the COMPILED reference (oracle/_ref = src/interpolation.c unmodified): its one-shot
    # mifi_interpolate_f and mifi_vector_reproject_values_f on the file's fields with fill values turned into NaN
    # (data2InterpolationArray, CDMInterpolator.cc:115-119) for the CLI's target axes 5,5.5,6,6.5 / 61.5,62,62.5
    nc = netcdf_file(os.path.join(ref, "test/hirlam12.nc"), "r", mmap=False)
    for name in ("Xc", "Yc", "geopotential_height", "air_potential_temperature", "x_wind", "y_wind"):
        out["hirlam_" + name] = native(nc.variables[name].data)
    out["hirlam_fill"] = np.float32(nc.variables["x_wind"]._FillValue)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
    from oracle import oracle as orc
    if orc.Reference.available():
        refc = orc.Reference()
        sphere = "+proj=latlong +a=6371000 +e=0 +no_defs"
        tx, ty = np.array([5, 5.5, 6, 6.5]), np.array([61.5, 62, 62.5])
        out["hirlam_target_x"], out["hirlam_target_y"] = tx, ty
        fields = {}
        for name in ("geopotential_height", "air_potential_temperature", "x_wind", "y_wind"):
            f = out["hirlam_" + name].astype(np.float32).reshape(4, 12, 17).copy()
            f[f == out["hirlam_fill"]] = np.nan
            rc, o = refc.interpolate_f(orc.BILINEAR, sphere, f, out["hirlam_Xc"], out["hirlam_Yc"], orc.LONGITUDE, orc.LATITUDE, 4, sphere,
                                       tx, ty, orc.LONGITUDE, orc.LATITUDE)
            assert rc == 1
            fields[name] = o
            out["hirlam_golden_" + name] = o
        rc, u, v = refc.vector_reproject_values(sphere, sphere, fields["x_wind"], fields["y_wind"], tx, ty, orc.LONGITUDE, orc.LATITUDE, 4)
        assert rc == 1
        out["hirlam_golden_x_wind_rotated"], out["hirlam_golden_y_wind_rotated"] = u.reshape(4, 3, 4), v.reshape(4, 3, 4)

    # test_merger (test/testMerger.cc:44-77): ga_2t_1 of test_merge_inner.nc merged into test_merge_outer.nc
    for tag, fn in (("inner", "test/test_merge_inner.nc"), ("outer", "test/test_merge_outer.nc")):
        nc = netcdf_file(os.path.join(ref, fn), "r", mmap=False)
        out[f"merge_{tag}_proj4"] = np.array(nc.variables["projection_regular_ll"].proj4.decode())
        for name in ("longitude", "latitude", "ga_2t_1"):
            out[f"merge_{tag}_{name}"] = native(nc.variables[name].data)

    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "interpolator_fixtures.npz")
    np.savez_compressed(dst, **out)
    for k, v in out.items():
        print(k, v.dtype, v.shape)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
