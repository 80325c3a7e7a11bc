// fimex_b200/csrc/proj.cuh -- fp64 map-projection core of the regridding path (K1).
//
// Replaces, on the GPU, what the reference obtains from PROJ.4 through pj_init_plus / pj_transform
// (call sites /root/reference/src/interpolation.c:355,396,644,700,773,1185,1233).  One thread transforms
// one point: inverse of the source CRS -> optional datum shift -> forward of the target CRS, with PROJ's
// "transient error => HUGE_VAL" convention, which mifi_points2position later maps to -999
// (interpolation.c:183-186).  Formulas: PROJ 4.9.x / Snyder, "Map Projections - A Working Manual";
// supported CRSs are those BASELINE.json's north_star names: lat/long, rotated pole (ob_tran over
// longlat), stereographic (polar, oblique, equatorial; sphere and ellipsoid) and Lambert conformal conic
// (1 or 2 standard parallels; sphere and ellipsoid).
//
// ProjDef is a plain struct: parsed on the host (proj_parse.cpp), passed to kernels by value.
#pragma once

#include <math.h>

#ifndef __CUDACC__
#define FB_HD
#else
#define FB_HD __host__ __device__ __forceinline__
#endif

namespace fb {

enum ProjKind { PK_LATLONG = 0, PK_OB_TRAN = 1, PK_STERE = 2, PK_LCC = 3 };
enum StereMode { SM_SOUTH = 0, SM_NORTH = 1, SM_OBLIQUE = 2, SM_EQUATOR = 3 };
enum DatumKind { DK_UNKNOWN = 0, DK_3PARAM = 1, DK_7PARAM = 2, DK_GRIDSHIFT = 3, DK_WGS84 = 4 };

struct ProjDef {
    int kind;
    int is_latlong;
    int over;
    int geoc;
    double a, ra, es, e, one_es, rone_es;
    double lam0, phi0, x0, y0, k0, to_meter, fr_meter;
    int datum_kind;
    double datum[7];
    double a_orig, es_orig;
    // stereographic
    int st_mode;
    double st_akm1, st_sin1, st_cos1; // sin/cos of phi0 (sphere) or of the conformal latitude X1 (ellipsoid)
    // Lambert conformal conic
    int lcc_ellips;
    double lcc_n, lcc_rho0, lcc_c;
    // rotated pole
    int ob_oblique;
    double ob_lamp, ob_cphip, ob_sphip;
};

// parse a proj4 string; returns 0 or a (negative) PROJ-style error number and fills `err`
int parse_proj(const char* definition, ProjDef* out, char* err, int errlen);

// whether a datum shift is required between two CRSs (pj_datum_transform's early exits)
bool needs_datum_shift(const ProjDef& s, const ProjDef& d);

#ifdef __CUDACC__

#define FB_HALFPI 1.5707963267948966
#define FB_FORTPI 0.78539816339744833
#define FB_ONEPI 3.14159265358979323846
#define FB_TWOPI 6.2831853071795864769
#define FB_SPI 3.14159265359
#define FB_HUGE (__longlong_as_double(0x7ff0000000000000LL)) /* HUGE_VAL */

// error numbers as PROJ uses them; transient ones become HUGE_VAL coordinates
#define FB_E_LATLON_LIMIT (-14)
#define FB_E_INVALID_XY (-15)
#define FB_E_ASIN_RANGE (-19)
#define FB_E_TOLERANCE (-20)
#define FB_E_PHI2 (-18)

__device__ __forceinline__ bool proj_err_is_transient(int e)
{
    return e == -14 || e == -15 || e == -17 || e == -20 || e == -27 || e == -45;
}

__device__ __forceinline__ double wrap_longitude(double lon)
{
    if (fabs(lon) <= FB_SPI)
        return lon;
    lon += FB_ONEPI;
    lon -= FB_TWOPI * floor(lon / FB_TWOPI);
    lon -= FB_ONEPI;
    return lon;
}

__device__ __forceinline__ double clamped_asin(double v, int& err)
{
    const double av = fabs(v);
    if (av >= 1.) {
        if (av > 1.00000000000001)
            err = FB_E_ASIN_RANGE;
        return v < 0. ? -FB_HALFPI : FB_HALFPI;
    }
    return asin(v);
}

__device__ __forceinline__ double guarded_atan2(double n, double d)
{
    return (fabs(n) < 1e-50 && fabs(d) < 1e-50) ? 0. : atan2(n, d);
}

// t = tan(pi/4 - phi/2) / ((1 - e sin phi)/(1 + e sin phi))^(e/2)   (Snyder 15-9)
__device__ __forceinline__ double conformal_t(double phi, double sinphi, double e)
{
    const double es = sinphi * e;
    return tan(.5 * (FB_HALFPI - phi)) / pow((1. - es) / (1. + es), .5 * e);
}

// latitude from t by fixed-point iteration (Snyder 7-9), at most 15 rounds, tolerance 1e-10
__device__ __forceinline__ double latitude_from_t(double ts, double e, int& err)
{
    const double half_e = .5 * e;
    double phi = FB_HALFPI - 2. * atan(ts);
    int rounds = 15;
    double step;
    do {
        const double con = e * sin(phi);
        step = FB_HALFPI - 2. * atan(ts * pow((1. - con) / (1. + con), half_e)) - phi;
        phi += step;
    } while (fabs(step) > 1.0e-10 && --rounds);
    if (rounds <= 0)
        err = FB_E_PHI2;
    return phi;
}

// ---------------------------------------------------------------------------------------------------
// stereographic
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stere_forward(const ProjDef& P, double lam, double phi, double& x, double& y, int& err)
{
    double sinlam, coslam;
    sincos(lam, &sinlam, &coslam);
    if (P.es == 0.) {
        double sinphi, cosphi;
        sincos(phi, &sinphi, &cosphi);
        if (P.st_mode == SM_EQUATOR || P.st_mode == SM_OBLIQUE) {
            double k = (P.st_mode == SM_EQUATOR) ? 1. + cosphi * coslam : 1. + P.st_sin1 * sinphi + P.st_cos1 * cosphi * coslam;
            if (k <= 1.e-10) {
                err = FB_E_TOLERANCE;
                return;
            }
            k = P.st_akm1 / k;
            x = k * cosphi * sinlam;
            y = k * ((P.st_mode == SM_EQUATOR) ? sinphi : P.st_cos1 * sinphi - P.st_sin1 * cosphi * coslam);
        } else {
            if (P.st_mode == SM_NORTH) {
                coslam = -coslam;
                phi = -phi;
            }
            if (fabs(phi - FB_HALFPI) < 1.e-8) {
                err = FB_E_TOLERANCE;
                return;
            }
            const double rho = P.st_akm1 * tan(FB_FORTPI + .5 * phi);
            x = sinlam * rho;
            y = rho * coslam;
        }
    } else {
        double sinphi = sin(phi);
        double rx = 0., ry = 0.;
        if (P.st_mode == SM_OBLIQUE || P.st_mode == SM_EQUATOR) {
            const double s = sinphi * P.e;
            const double chi = 2. * atan(tan(.5 * (FB_HALFPI + phi)) * pow((1. - s) / (1. + s), .5 * P.e)) - FB_HALFPI;
            double sinX, cosX;
            sincos(chi, &sinX, &cosX);
            if (P.st_mode == SM_OBLIQUE) {
                const double A = P.st_akm1 / (P.st_cos1 * (1. + P.st_sin1 * sinX + P.st_cos1 * cosX * coslam));
                ry = A * (P.st_cos1 * sinX - P.st_sin1 * cosX * coslam);
                rx = A * cosX;
            } else {
                const double A = 2. * P.st_akm1 / (1. + cosX * coslam);
                ry = A * sinX;
                rx = A * cosX;
            }
        } else {
            if (P.st_mode == SM_SOUTH) {
                phi = -phi;
                coslam = -coslam;
                sinphi = -sinphi;
            }
            rx = P.st_akm1 * conformal_t(phi, sinphi, P.e);
            ry = -rx * coslam;
        }
        x = rx * sinlam;
        y = ry;
    }
}

__device__ __forceinline__ void stere_inverse(const ProjDef& P, double x, double y, double& lam, double& phi, int& err)
{
    if (P.es == 0.) {
        const double rh = hypot(x, y);
        double c = 2. * atan(rh / P.st_akm1);
        double sinc, cosc;
        sincos(c, &sinc, &cosc);
        lam = 0.;
        switch (P.st_mode) {
        case SM_EQUATOR:
            phi = (fabs(rh) <= 1.e-10) ? 0. : asin(y * sinc / rh);
            if (cosc != 0. || x != 0.)
                lam = atan2(x * sinc, cosc * rh);
            break;
        case SM_OBLIQUE:
            phi = (fabs(rh) <= 1.e-10) ? P.phi0 : asin(cosc * P.st_sin1 + y * sinc * P.st_cos1 / rh);
            c = cosc - P.st_sin1 * sin(phi);
            if (c != 0. || x != 0.)
                lam = atan2(x * sinc * P.st_cos1, c * rh);
            break;
        default:
            if (P.st_mode == SM_NORTH)
                y = -y;
            phi = (fabs(rh) <= 1.e-10) ? P.phi0 : asin(P.st_mode == SM_SOUTH ? -cosc : cosc);
            lam = (x == 0. && y == 0.) ? 0. : atan2(x, y);
            break;
        }
    } else {
        double tp, phi_l, halfe, halfpi;
        const double rho = hypot(x, y);
        if (P.st_mode == SM_OBLIQUE || P.st_mode == SM_EQUATOR) {
            tp = 2. * atan2(rho * P.st_cos1, P.st_akm1);
            double sinphi, cosphi;
            sincos(tp, &sinphi, &cosphi);
            phi_l = (rho == 0.0) ? asin(cosphi * P.st_sin1) : asin(cosphi * P.st_sin1 + (y * sinphi * P.st_cos1 / rho));
            tp = tan(.5 * (FB_HALFPI + phi_l));
            x *= sinphi;
            y = rho * P.st_cos1 * cosphi - y * P.st_sin1 * sinphi;
            halfpi = FB_HALFPI;
            halfe = .5 * P.e;
        } else {
            if (P.st_mode == SM_NORTH)
                y = -y;
            tp = -rho / P.st_akm1;
            phi_l = FB_HALFPI - 2. * atan(tp);
            halfpi = -FB_HALFPI;
            halfe = -.5 * P.e;
        }
        for (int i = 0; i < 8; ++i) {
            const double s = P.e * sin(phi_l);
            const double p = 2. * atan(tp * pow((1. + s) / (1. - s), halfe)) - halfpi;
            if (fabs(phi_l - p) < 1.e-10) {
                phi = (P.st_mode == SM_SOUTH) ? -p : p;
                lam = (x == 0. && y == 0.) ? 0. : atan2(x, y);
                return;
            }
            phi_l = p;
        }
        err = FB_E_TOLERANCE;
    }
}

// ---------------------------------------------------------------------------------------------------
// Lambert conformal conic
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lcc_forward(const ProjDef& P, double lam, double phi, double& x, double& y, int& err)
{
    double rho;
    if (fabs(fabs(phi) - FB_HALFPI) < 1.e-10) {
        if ((phi * P.lcc_n) <= 0.) {
            err = FB_E_TOLERANCE;
            return;
        }
        rho = 0.;
    } else {
        rho = P.lcc_c * (P.lcc_ellips ? pow(conformal_t(phi, sin(phi), P.e), P.lcc_n) : pow(tan(FB_FORTPI + .5 * phi), -P.lcc_n));
    }
    double s, c;
    sincos(lam * P.lcc_n, &s, &c);
    x = P.k0 * (rho * s);
    y = P.k0 * (P.lcc_rho0 - rho * c);
}

__device__ __forceinline__ void lcc_inverse(const ProjDef& P, double x, double y, double& lam, double& phi, int& err)
{
    x /= P.k0;
    y /= P.k0;
    y = P.lcc_rho0 - y;
    double rho = hypot(x, y);
    if (rho != 0.0) {
        if (P.lcc_n < 0.) {
            rho = -rho;
            x = -x;
            y = -y;
        }
        if (P.lcc_ellips) {
            phi = latitude_from_t(pow(rho / P.lcc_c, 1. / P.lcc_n), P.e, err);
        } else {
            phi = 2. * atan(pow(P.lcc_c / rho, 1. / P.lcc_n)) - FB_HALFPI;
        }
        lam = atan2(x, y) / P.lcc_n;
    } else {
        lam = 0.;
        phi = P.lcc_n > 0. ? FB_HALFPI : -FB_HALFPI;
    }
}

// ---------------------------------------------------------------------------------------------------
// rotated pole: ob_tran with a lat/long link ("+o_lat_p [+o_lon_p]" form).  The link maps (lam, phi) to
// (lam/a, phi/a) and back, which the generic scaling below undoes -- the divisions are kept so that the
// results round the same way as the CPU library's.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rotpole_forward(const ProjDef& P, double lam, double phi, double& x, double& y, int& err)
{
    double sinlam, coslam, sinphi, cosphi;
    sincos(lam, &sinlam, &coslam);
    sincos(phi, &sinphi, &cosphi);
    double rl, rp;
    if (P.ob_oblique) {
        rl = wrap_longitude(guarded_atan2(cosphi * sinlam, P.ob_sphip * cosphi * coslam + P.ob_cphip * sinphi) + P.ob_lamp);
        rp = clamped_asin(P.ob_sphip * sinphi - P.ob_cphip * cosphi * coslam, err);
    } else {
        rl = wrap_longitude(guarded_atan2(cosphi * sinlam, sinphi) + P.ob_lamp);
        rp = clamped_asin(-cosphi * coslam, err);
    }
    x = rl / P.a;
    y = rp / P.a;
}

__device__ __forceinline__ void rotpole_inverse(const ProjDef& P, double x, double y, double& lam, double& phi, int& err)
{
    double rp = y * P.a;
    double rl = x * P.a;
    if (rl == FB_HUGE) {
        lam = rl;
        phi = rp;
        return;
    }
    double sinphi, cosphi;
    sincos(rp, &sinphi, &cosphi);
    if (P.ob_oblique) {
        rl -= P.ob_lamp;
        double sinl, cosl;
        sincos(rl, &sinl, &cosl);
        phi = clamped_asin(P.ob_sphip * sinphi + P.ob_cphip * cosphi * cosl, err);
        lam = guarded_atan2(cosphi * sinl, P.ob_sphip * cosphi * cosl - P.ob_cphip * sinphi);
    } else {
        const double t = rl - P.ob_lamp;
        double sint, cost;
        sincos(t, &sint, &cost);
        lam = guarded_atan2(cosphi * sint, -sinphi);
        phi = clamped_asin(cosphi * cost, err);
    }
}

// ---------------------------------------------------------------------------------------------------
// generic wrappers (what pj_fwd / pj_inv add around a projection)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int projected_from_geodetic(const ProjDef& P, double lam, double phi, double& x, double& y)
{
    const double t = fabs(phi) - FB_HALFPI;
    if (t > 1.0e-12 || fabs(lam) > 10.) {
        x = y = FB_HUGE;
        return FB_E_LATLON_LIMIT;
    }
    if (fabs(t) <= 1.0e-12)
        phi = phi < 0. ? -FB_HALFPI : FB_HALFPI;
    else if (P.geoc)
        phi = atan(P.rone_es * tan(phi));
    lam -= P.lam0;
    if (!P.over)
        lam = wrap_longitude(lam);
    int err = 0;
    double px = 0., py = 0.;
    switch (P.kind) {
    case PK_OB_TRAN:
        rotpole_forward(P, lam, phi, px, py, err);
        break;
    case PK_STERE:
        stere_forward(P, lam, phi, px, py, err);
        break;
    case PK_LCC:
        lcc_forward(P, lam, phi, px, py, err);
        break;
    default:
        px = lam / P.a;
        py = phi / P.a;
        break;
    }
    if (err) {
        x = y = FB_HUGE;
        return err;
    }
    x = P.fr_meter * (P.a * px + P.x0);
    y = P.fr_meter * (P.a * py + P.y0);
    return 0;
}

__device__ __forceinline__ int geodetic_from_projected(const ProjDef& P, double x, double y, double& lam, double& phi)
{
    if (x == FB_HUGE || y == FB_HUGE) {
        lam = phi = FB_HUGE;
        return FB_E_INVALID_XY;
    }
    x = (x * P.to_meter - P.x0) * P.ra;
    y = (y * P.to_meter - P.y0) * P.ra;
    int err = 0;
    double l = 0., p = 0.;
    switch (P.kind) {
    case PK_OB_TRAN:
        rotpole_inverse(P, x, y, l, p, err);
        break;
    case PK_STERE:
        stere_inverse(P, x, y, l, p, err);
        break;
    case PK_LCC:
        lcc_inverse(P, x, y, l, p, err);
        break;
    default:
        p = y * P.a;
        l = x * P.a;
        break;
    }
    if (err) {
        lam = phi = FB_HUGE;
        return err;
    }
    l += P.lam0;
    if (!P.over)
        l = wrap_longitude(l);
    if (P.geoc && fabs(fabs(p) - FB_HALFPI) > 1.0e-12)
        p = atan(P.one_es * tan(p));
    lam = l;
    phi = p;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// datum shift through geocentric coordinates (3- and 7-parameter), height 0 in, height discarded out
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void datum_shift(const ProjDef& S, const ProjDef& D, double& lon, double& lat)
{
    double h = 0.;
    // geodetic -> geocentric on the source ellipsoid
    if (lat < -FB_HALFPI && lat > -1.001 * FB_HALFPI)
        lat = -FB_HALFPI;
    else if (lat > FB_HALFPI && lat < 1.001 * FB_HALFPI)
        lat = FB_HALFPI;
    else if (lat < -FB_HALFPI || lat > FB_HALFPI) {
        lon = lat = FB_HUGE;
        return;
    }
    if (lon > FB_ONEPI)
        lon -= (2 * FB_ONEPI);
    double sin_lat, cos_lat, sin_lon, cos_lon;
    sincos(lat, &sin_lat, &cos_lat);
    sincos(lon, &sin_lon, &cos_lon);
    const double rn = S.a_orig / sqrt(1.0 - S.es_orig * sin_lat * sin_lat);
    double X = (rn + h) * cos_lat * cos_lon;
    double Y = (rn + h) * cos_lat * sin_lon;
    double Z = ((rn * (1 - S.es_orig)) + h) * sin_lat;
    // source datum -> WGS84 -> target datum
    if (S.datum_kind == DK_3PARAM) {
        X += S.datum[0];
        Y += S.datum[1];
        Z += S.datum[2];
    } else if (S.datum_kind == DK_7PARAM) {
        const double* q = S.datum;
        const double xo = q[6] * (X - q[5] * Y + q[4] * Z) + q[0];
        const double yo = q[6] * (q[5] * X + Y - q[3] * Z) + q[1];
        const double zo = q[6] * (-q[4] * X + q[3] * Y + Z) + q[2];
        X = xo;
        Y = yo;
        Z = zo;
    }
    if (D.datum_kind == DK_3PARAM) {
        X -= D.datum[0];
        Y -= D.datum[1];
        Z -= D.datum[2];
    } else if (D.datum_kind == DK_7PARAM) {
        const double* q = D.datum;
        const double xt = (X - q[0]) / q[6];
        const double yt = (Y - q[1]) / q[6];
        const double zt = (Z - q[2]) / q[6];
        X = xt + q[5] * yt - q[4] * zt;
        Y = -q[5] * xt + yt + q[3] * zt;
        Z = q[4] * xt - q[3] * yt + zt;
    }
    // geocentric -> geodetic on the target ellipsoid (iterative, tolerance 1e-12, at most 30 rounds)
    const double a = D.a_orig, es = D.es_orig;
    const double P = sqrt(X * X + Y * Y);
    const double RR = sqrt(X * X + Y * Y + Z * Z);
    if (P / a < 1.E-12) {
        lon = 0.;
        if (RR / a < 1.E-12) {
            lat = FB_HALFPI;
            return;
        }
    } else {
        lon = atan2(Y, X);
    }
    const double CT = Z / RR, ST = P / RR;
    double RX = 1.0 / sqrt(1.0 - es * (2.0 - es) * ST * ST);
    double CPHI0 = ST * (1.0 - es) * RX;
    double SPHI0 = CT * RX;
    double CPHI, SPHI, SDPHI;
    int iter = 0;
    do {
        iter++;
        const double RN = a / sqrt(1.0 - es * SPHI0 * SPHI0);
        h = P * CPHI0 + Z * SPHI0 - RN * (1.0 - es * SPHI0 * SPHI0);
        const double RK = es * RN / (RN + h);
        RX = 1.0 / sqrt(1.0 - RK * (2.0 - RK) * ST * ST);
        CPHI = ST * (1.0 - RK) * RX;
        SPHI = CT * RX;
        SDPHI = SPHI * CPHI0 - CPHI * SPHI0;
        CPHI0 = CPHI;
        SPHI0 = SPHI;
    } while (SDPHI * SDPHI > 1.E-24 && iter < 30);
    lat = atan(SPHI / fabs(CPHI));
}

// One point through the whole pipeline.  Returns 0, or a non-transient error number that makes the whole
// call fail (the reference then returns MIFI_ERROR); transient failures leave HUGE_VAL in x and y.
__device__ __forceinline__ int transform_point(const ProjDef& S, const ProjDef& D, bool shift, bool single_point, double& x, double& y)
{
    if (!S.is_latlong) {
        if (x != FB_HUGE) {
            double lam, phi;
            const int e = geodetic_from_projected(S, x, y, lam, phi);
            if (e != 0) {
                if (single_point || !proj_err_is_transient(e))
                    return e;
                lam = phi = FB_HUGE;
            }
            x = lam;
            y = phi;
        }
    }
    if (shift && x != FB_HUGE)
        datum_shift(S, D, x, y);
    if (!D.is_latlong) {
        if (x != FB_HUGE) {
            double px, py;
            const int e = projected_from_geodetic(D, x, y, px, py);
            if (e != 0) {
                if (single_point || !proj_err_is_transient(e))
                    return e;
                px = py = FB_HUGE;
            }
            x = px;
            y = py;
        }
    }
    return 0;
}

#endif // __CUDACC__

} // namespace fb
