"""Independent pins for the fp64 coordinate transforms (PROJ.4 is a third-party dependency that is not under
/root/reference, so `oracle/pj_oracle.c` and the GPU `proj.cuh` are both restatements).  Nothing in this module comes from
either of them:

  * PUBLISHED numerical examples of J. P. Snyder, "Map Projections -- A Working Manual" (USGS Prof. Paper 1395, 1987),
    Appendix A: the printed inputs and the printed results, to the printed digits;
  * the same projections evaluated from Snyder's equations written out here in numpy (equation numbers cited), on random
    points -- a third, independent implementation;
  * rotated pole (PROJ `ob_tran +o_proj=longlat`) as an explicit 3-D change of basis built from the CF definition of the
    rotated pole (grid_north_pole_latitude / _longitude; the reference maps those to `+o_lat_p` and `+lon_0 = pole_lon - 180`
    in src/coordSys/RotatedLatitudeLongitudeProjection.cc), with no spherical-trigonometry formula shared with PROJ.

tests/test_proj_known_answers.py asserts them against the CPU oracle (`-m "not gpu"`) and against the GPU through
`mifi_project_values` (`-m gpu`).  What they replace in the reference: src/interpolation.c:1158-1244 -> pj_transform.
"""
from __future__ import annotations

import numpy as np

DEG = np.pi / 180.0

# ------------------------------------------------------------------------------------------------------------------
# Snyder's printed examples: (name, proj string, lon deg, lat deg, x, y, absolute tolerance = half a unit of the last printed digit)
# ------------------------------------------------------------------------------------------------------------------
SNYDER = [
    # p. 295-296, Lambert Conformal Conic, sphere R = 1: std parallels 33 / 45 N, origin 23 N 96 W; point 35 N 75 W
    ("lcc sphere", "+proj=lcc +lat_1=33 +lat_2=45 +lat_0=23 +lon_0=-96 +R=1 +no_defs", -75.0, 35.0, 0.2966785, 0.2462112, 6e-8),
    # p. 296-297, same on the Clarke 1866 ellipsoid (a = 6378206.4 m, e^2 = 0.00676866)
    ("lcc clarke 1866", "+proj=lcc +lat_1=33 +lat_2=45 +lat_0=23 +lon_0=-96 +a=6378206.4 +es=0.00676866 +no_defs", -75.0, 35.0,
     1894410.9, 1564649.5, 0.06),
    # p. 312-313, Stereographic, sphere R = 1, oblique: centre 40 N 100 W, k0 = 1; point 30 N 75 W
    ("stere oblique sphere", "+proj=stere +lat_0=40 +lon_0=-100 +k=1 +R=1 +no_defs", -75.0, 30.0, 0.3807224, -0.1263802, 6e-8),
    # p. 315, Stereographic, International ellipsoid (a = 6378388 m, e^2 = 0.00672267), south polar aspect, true scale at 71 S,
    # central meridian 100 W; point 75 S 150 E
    ("stere polar ellipsoid", "+proj=stere +lat_0=-90 +lat_ts=-71 +lon_0=-100 +a=6378388 +es=0.00672267 +no_defs", 150.0, -75.0,
     -1540033.6, -560526.4, 0.06),
]

LATLONG_FOR = {  # the geographic CRS on the same figure (no datum on either side => no shift, SURVEY.md 8c')
    "lcc sphere": "+proj=latlong +R=1 +no_defs",
    "lcc clarke 1866": "+proj=latlong +a=6378206.4 +es=0.00676866 +no_defs",
    "stere oblique sphere": "+proj=latlong +R=1 +no_defs",
    "stere polar ellipsoid": "+proj=latlong +a=6378388 +es=0.00672267 +no_defs",
}


# ------------------------------------------------------------------------------------------------------------------
# Snyder's equations, written out
# ------------------------------------------------------------------------------------------------------------------
def _t(phi, e):
    """eq. 15-9: t = tan(pi/4 - phi/2) / [(1 - e sin phi) / (1 + e sin phi)]^(e/2)"""
    s = np.sin(phi)
    return np.tan(np.pi / 4 - phi / 2) / ((1 - e * s) / (1 + e * s))**(e / 2)


def _m(phi, e):
    """eq. 14-15: m = cos phi / (1 - e^2 sin^2 phi)^(1/2)"""
    return np.cos(phi) / np.sqrt(1 - (e * np.sin(phi))**2)


def lcc_forward(lon, lat, lat1, lat2, lat0, lon0, a, es):
    """Lambert Conformal Conic, ellipsoid (sphere for es = 0): eqs. 15-8, 15-10, 15-7a, 14-4, 14-1, 14-2"""
    e = np.sqrt(es)
    m1, m2 = _m(lat1, e), _m(lat2, e)
    t0, t1, t2, t = _t(lat0, e), _t(lat1, e), _t(lat2, e), _t(lat, e)
    n = np.sin(lat1) if abs(lat1 - lat2) < 1e-12 else (np.log(m1) - np.log(m2)) / (np.log(t1) - np.log(t2))  # 15-8
    F = m1 / (n * t1**n)  # 15-10
    rho0 = a * F * t0**n  # 15-7a
    rho = a * F * t**n  # 15-7
    theta = n * (lon - lon0)  # 14-4
    return rho * np.sin(theta), rho0 - rho * np.cos(theta)  # 14-1, 14-2


def stere_polar_forward(lon, lat, lat_ts, lon0, a, es, south):
    """Polar stereographic, ellipsoid: eqs. 21-34 (rho = a m_c t / t_c), 21-30, 21-31; the south polar aspect is the north
    polar one with the signs of lat, lat_ts, lon, lon0 and of the results x, y reversed (Snyder p. 161)"""
    if south:
        x, y = stere_polar_forward(-lon, -lat, -lat_ts, -lon0, a, es, False)
        return -x, -y
    e = np.sqrt(es)
    rho = a * _m(lat_ts, e) * _t(lat, e) / _t(lat_ts, e)
    return rho * np.sin(lon - lon0), -rho * np.cos(lon - lon0)


def stere_oblique_sphere_forward(lon, lat, lat0, lon0, R, k0):
    """Stereographic, sphere, oblique aspect: eqs. 21-4, 21-2, 21-3"""
    dl = lon - lon0
    k = 2 * k0 / (1 + np.sin(lat0) * np.sin(lat) + np.cos(lat0) * np.cos(lat) * np.cos(dl))
    return R * k * np.cos(lat) * np.sin(dl), R * k * (np.cos(lat0) * np.sin(lat) - np.sin(lat0) * np.cos(lat) * np.cos(dl))


# ------------------------------------------------------------------------------------------------------------------
# rotated pole as a change of basis
# ------------------------------------------------------------------------------------------------------------------
def _unit(lon, lat):
    return np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], axis=-1)


def rotated_pole_basis(o_lat_p, lon_0):
    """Rows = the rotated frame's x', y', z' axes in geographic cartesian coordinates.  CF: the rotated north pole sits at
    geographic (grid_north_pole_latitude, grid_north_pole_longitude) = (o_lat_p, lon_0 + 180 deg); the rotated origin
    (0, 0) lies on the geographic meridian lon_0, a quarter circle from that pole on the side away from it."""
    ez = _unit(lon_0 + np.pi, o_lat_p)
    ex = _unit(lon_0, np.pi / 2 - o_lat_p)
    ey = np.cross(ez, ex)
    return np.stack([ex, ey, ez])


def geographic_to_rotated(lon, lat, o_lat_p, lon_0):
    B = rotated_pole_basis(o_lat_p, lon_0)
    q = _unit(np.asarray(lon), np.asarray(lat)) @ B.T
    return np.arctan2(q[..., 1], q[..., 0]), np.arcsin(np.clip(q[..., 2], -1, 1))


def rotated_to_geographic(rlon, rlat, o_lat_p, lon_0):
    B = rotated_pole_basis(o_lat_p, lon_0)
    p = _unit(np.asarray(rlon), np.asarray(rlat)) @ B
    return np.arctan2(p[..., 1], p[..., 0]), np.arcsin(np.clip(p[..., 2], -1, 1))


def angle_diff(a, b):
    d = np.abs(np.asarray(a) - np.asarray(b)) % (2 * np.pi)
    return np.minimum(d, 2 * np.pi - d)
