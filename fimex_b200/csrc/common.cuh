// fimex_b200/csrc/common.cuh -- shared declarations of the B200 regridding library (sm_100a only).
//
// Internal header: the public C ABI is include/fimex_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#define FB_OK 1       // MIFI_OK,    reference include/fimex/mifi_constants.h:261
#define FB_ERROR (-1) // MIFI_ERROR, reference include/fimex/mifi_constants.h:259
#define FB_PI 3.1415926535897932384626433832795 // MIFI_PI, mifi_constants.h:42
#define FB_DEG_TO_RAD .0174532925199432958      // proj_api.h DEG_TO_RAD, used by convertAxis (interpolation.c:225)
#define FB_RAD_TO_DEG 57.29577951308232

#define FB_AXIS_PROJ 0
#define FB_AXIS_LONGITUDE 1
#define FB_AXIS_LATITUDE 2

// interpolation methods, same numbering as enum mifi_interpol_method (mifi_constants.h:52-147)
enum {
    FB_NN = 0,
    FB_BILINEAR,
    FB_BICUBIC,
    FB_COORD_NN,
    FB_COORD_NN_KD,
    FB_FWD_SUM,
    FB_FWD_MEAN,
    FB_FWD_MEDIAN,
    FB_FWD_MAX,
    FB_FWD_MIN,
    FB_FWD_UNDEF_SUM,
    FB_FWD_UNDEF_MEAN,
    FB_FWD_UNDEF_MEDIAN,
    FB_FWD_UNDEF_MAX,
    FB_FWD_UNDEF_MIN
};

// bilinear table modes: which taps of the 2x2 cell at `off` are used and with which formula
// (reference src/interpolation.c:889-953)
enum {
    FB_BL_FULL = 0, // :890-902  4 taps
    FB_BL_XLIN = 1, // :904-913  linear in x on the row at off      (taps off, off+1)
    FB_BL_YLIN = 2, // :925-933  linear in y on the column at off   (taps off, off+inX)
    FB_BL_NEAR = 3, // :935-942  single tap at off
    FB_BL_NAN = 4   // :916-918, :944-946, :950-952 and the undefined read at :936 (y0 == iy)
};

namespace fb {

// last error text of the calling thread (fb200_last_error)
void set_error(const std::string& msg);
const char* last_error();

#define FB_CUDA_CHECK(expr)                                                                                  \
    do {                                                                                                     \
        cudaError_t fb_e_ = (expr);                                                                          \
        if (fb_e_ != cudaSuccess) {                                                                          \
            ::fb::set_error(std::string(#expr) + ": " + cudaGetErrorString(fb_e_));                          \
            return FB_ERROR;                                                                                 \
        }                                                                                                    \
    } while (0)

#define FB_REQUIRE(cond, msg)                                                                                \
    do {                                                                                                     \
        if (!(cond)) {                                                                                       \
            ::fb::set_error(msg);                                                                            \
            return FB_ERROR;                                                                                 \
        }                                                                                                    \
    } while (0)

// canonical quiet NaN 0x7fc00000 == MIFI_UNDEFINED_F (nanf(""), mifi_constants.h:254)
__host__ __device__ inline float undef_f()
{
#ifdef __CUDA_ARCH__
    return __int_as_float(0x7fc00000);
#else
    union {
        uint32_t u;
        float f;
    } v;
    v.u = 0x7fc00000u;
    return v.f;
#endif
}

inline int ceil_div(long long a, long long b)
{
    return (int)((a + b - 1) / b);
}

int sm_count(); // multiprocessors of the current device (148 on B200)

// launch counter: every kernel launch of this library bumps it (bench.py reports it as gpu_launches)
void count_launch(int n = 1);
unsigned long long launches();

} // namespace fb
