// fimex_b200/csrc/convert.cuh -- K10: the type / fill-value adapters on either side of the gather, as device functors
// that the gather kernels apply while they load and store.
//
// Reference (/root/reference):
//   data2InterpolationArray  src/CDMInterpolator.cc:115-119 = Data::asFloat() (static_cast per element,
//                            src/DataImpl.h:385-389 with include/fimex/Utils.h:88-113) + mifi_bad2nanf
//                            (src/interpolation.c:1775-1783: value == (float)badValue -> MIFI_UNDEFINED_F; disabled when
//                            badValue is NaN)
//   interpolationArray2Data  src/CDMInterpolator.cc:121-124 = convertDataType(MIFI_UNDEFINED_F, 1., 0., type, badValue,
//                            1., 0.) -> ScaleValue<float, OUT> (include/fimex/Utils.h:444-464):
//                              NaN -> static_cast<OUT>(badValue); otherwise data_caster<OUT, double>(1.*in + 0.), which is
//                              static_cast<OUT>(int(lround(d))) for integer OUT (Utils.h:72-75, 98-99: MetNoFimex::round
//                              returns int) and static_cast<OUT>(d) for float / double.  The "+ 0." turns -0 into +0.
#pragma once

#ifndef FB_STORE_ROUND_MODE
#define FB_STORE_ROUND_MODE 1
#endif

#include "common.cuh"

namespace fb {

// CDMDataType, include/fimex/CDMDataType.h:35-49 (same numbering)
enum {
    FB_T_NAT = 0,
    FB_T_CHAR,
    FB_T_SHORT,
    FB_T_INT,
    FB_T_FLOAT,
    FB_T_DOUBLE,
    FB_T_STRING,
    FB_T_UCHAR,
    FB_T_USHORT,
    FB_T_UINT,
    FB_T_INT64,
    FB_T_UINT64
};

inline size_t type_size(int t)
{
    switch (t) {
    case FB_T_CHAR:
    case FB_T_UCHAR:
        return 1;
    case FB_T_SHORT:
    case FB_T_USHORT:
        return 2;
    case FB_T_INT:
    case FB_T_UINT:
    case FB_T_FLOAT:
        return 4;
    case FB_T_DOUBLE:
    case FB_T_INT64:
    case FB_T_UINT64:
        return 8;
    default:
        return 0;
    }
}

template <class T>
struct is_fp {
    static constexpr bool value = false;
};
template <>
struct is_fp<float> {
    static constexpr bool value = true;
};
template <>
struct is_fp<double> {
    static constexpr bool value = true;
};

// The value the gather produced, stored as is: CachedInterpolationInterface::interpolateValues' own output
// (NaN = MIFI_UNDEFINED_F stays NaN)
struct StorePlain {
    typedef float type;
    __device__ __forceinline__ float operator()(float v) const { return v; }
};

// interpolationArray2Data for one value: ScaleValue<float, OUT>(NaN, 1, 0, badValue, 1, 0)
template <class OUT>
struct StoreAs {
    typedef OUT type;
    OUT fill; // static_cast<OUT>(badValue), converted on the host
    __device__ __forceinline__ OUT operator()(float v) const
    {
        if (sizeof(OUT) == 4 && is_fp<OUT>::value)
            return isnan(v) ? fill : (OUT)__fadd_rn(v, 0.f); // == (float)(1.*in + 0.): only -0 changes (to +0)
        if (is_fp<OUT>::value)
            return isnan(v) ? fill : (OUT)__dadd_rn((double)v, 0.); // 1.*in + 0. in fp64
        // lround (half away from zero) in long, narrowed to int by MetNoFimex::round, then to OUT.  Rounded TOWARDS ZERO,
        // v + copysign(0.5, v) never crosses an integer (wherever the sum is inexact the integers are representable), so
        // truncating it IS lround(v) for every finite float: LOP3, FADD.RZ, F2I.S64 -- one conversion-pipe instruction and no
        // fp64 (the earlier form, trunc((double)v + copysign(0.5, v)), needed F2F + DADD + F2I.S64.F64: int16 output of config 2
        // 14.9 ms -> 13.3 ms).  Exhaustively checked per exponent in tests/test_gpu_parity.py.  Values beyond the range of long
        // are undefined in the reference.
#if FB_STORE_ROUND_MODE == 1
        const float half = __uint_as_float(0x3F000000u | (__float_as_uint(v) & 0x80000000u));
        const long long r = __float2ll_rz(__fadd_rz(v, half));
        return isnan(v) ? fill : (OUT)(int)r;
#else // A/B: the fp64 form
        const double half = __hiloint2double(0x3FE00000 | (int)(__float_as_uint(v) & 0x80000000u), 0);
        const long long r = __double2ll_rz(__dadd_rn((double)v, half));
        return isnan(v) ? fill : (OUT)(int)r;
#endif
    }
};

// asFloat() + mifi_bad2nanf for one value
template <class IN>
__device__ __forceinline__ float load_as_float(IN x, bool has_bad, float bad)
{
    const float f = (float)x; // static_cast<float>(in); round-to-nearest for double and 64-bit integers like x86-64
    return (has_bad && f == bad) ? undef_f() : f;
}

// four consecutive converted values with one store instruction (dst must be aligned to 4 * sizeof(T))
template <class T>
__device__ __forceinline__ void store_vec4(T* dst, T a, T b, T c, T d)
{
    if constexpr (sizeof(T) == 1) {
        union {
            T t[4];
            unsigned u;
        } w;
        w.t[0] = a, w.t[1] = b, w.t[2] = c, w.t[3] = d;
        __stcs(reinterpret_cast<unsigned*>(dst), w.u);
    } else if constexpr (sizeof(T) == 2) {
        union {
            T t[4];
            uint2 u;
        } w;
        w.t[0] = a, w.t[1] = b, w.t[2] = c, w.t[3] = d;
        __stcs(reinterpret_cast<uint2*>(dst), w.u);
    } else if constexpr (sizeof(T) == 4) {
        union {
            T t[4];
            uint4 u;
        } w;
        w.t[0] = a, w.t[1] = b, w.t[2] = c, w.t[3] = d;
        __stcs(reinterpret_cast<uint4*>(dst), w.u);
    } else {
        union {
            T t[4];
            uint4 u[2];
        } w;
        w.t[0] = a, w.t[1] = b, w.t[2] = c, w.t[3] = d;
        __stcs(reinterpret_cast<uint4*>(dst), w.u[0]);
        __stcs(reinterpret_cast<uint4*>(dst) + 1, w.u[1]);
    }
}

// what a slice call asks of the kernels around the gather
struct SliceConv {
    bool fill_in = false;  // replace values equal to bad_in[f] by NaN while loading (mifi_bad2nanf)
    float bad_in[2] = {0.f, 0.f};
    bool convert_out = false; // apply interpolationArray2Data while storing; false: plain float output
    int out_type = FB_T_FLOAT;
    double fill_out = 0.;
};

// host-side static_cast<OUT>(badValue) with defined behaviour for the out-of-range cases a sane fill value never hits
template <class OUT>
inline OUT cast_fill(double v)
{
    return static_cast<OUT>(v);
}

} // namespace fb
