// fimex_b200/csrc/tables.cuh -- per-point classification shared by the table compilers.
#pragma once

#include "common.cuh"

namespace fb {

// Positions whose magnitude does not fit an int are "outside": the reference narrows lround()/floor() to
// int (interpolation.c:864-865, 883-886), which is undefined for such values; finite positions of real
// grids never get there (-999 marks projection failures).
__device__ __forceinline__ bool fits_int(double v)
{
    return v > -2147483000.0 && v < 2147483000.0;
}


// One target point of a bilinear table: which source cell it reads and with which formula
// (reference src/interpolation.c:881-957).  Returns {off, bits(xfrac), bits(yfrac), mode}.
__device__ __forceinline__ int4 classify_bilinear(double x, double y, int ix, int iy)
{
    int4 e = make_int4(0, 0, 0, FB_BL_NAN);
    if (fits_int(x) && fits_int(y)) {
        const int x0 = (int)floor(x), y0 = (int)floor(y);
        const float xf = __double2float_rn(__dsub_rn(x, (double)x0)); // :885
        const float yf = __double2float_rn(__dsub_rn(y, (double)y0)); // :888
        e.y = __float_as_int(xf);
        e.z = __float_as_int(yf);
        const bool x_in = (0 <= x0) && (x0 + 1 < ix);
        const bool y_in = (0 <= y0) && (y0 + 1 < iy);
        if (x_in && y_in) {
            e.x = y0 * ix + x0;
            e.w = FB_BL_FULL;
        } else if (x_in) {
            const long long ry = llround(y);
            if (ry >= 0 && ry < iy) {
                e.x = (int)ry * ix + x0;
                e.w = FB_BL_XLIN;
            }
        } else {
            const long long rx = llround(x);
            if (rx >= 0 && rx < ix) {
                if (y_in) {
                    e.x = y0 * ix + (int)rx;
                    e.w = FB_BL_YLIN;
                } else {
                    const long long ry = llround(y);
                    // ry == iy is the reference's out-of-bounds read (:936); NaN here
                    if (ry >= 0 && ry < iy) {
                        e.x = (int)ry * ix + (int)rx;
                        e.w = FB_BL_NEAR;
                    }
                }
            }
        }
    }
    return e;
}

// Nearest neighbour as a one-tap entry of the same table: {off, 0, 0, NEAR or NAN}; lround semantics of
// src/interpolation.c:864-868 (half away from zero)
__device__ __forceinline__ int4 classify_nn(double x, double y, int ix, int iy)
{
    int4 e = make_int4(0, 0, 0, FB_BL_NAN);
    if (fits_int(x) && fits_int(y)) {
        const long long rx = llround(x), ry = llround(y);
        if (rx >= 0 && rx < ix && ry >= 0 && ry < iy) {
            e.x = (int)(ry * ix + rx);
            e.w = FB_BL_NEAR;
        }
    }
    return e;
}

} // namespace fb
