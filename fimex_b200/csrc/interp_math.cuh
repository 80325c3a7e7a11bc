// fimex_b200/csrc/interp_math.cuh -- the per-point arithmetic of the gathers, shared by the direct kernels
// (gather_kernels.cu) and the staged ones (staged_kernels.cu, bicubic_staged.cu).
//
// Every function replays the reference's operation order (/root/reference/src/interpolation.c) with explicit
// round-to-nearest intrinsics so that nvcc cannot contract a*b+c into an FMA: the results are bit-identical to the
// reference built for x86-64 (SSE2, no FMA; SURVEY.md 8a trap 7).
#pragma once

#include "common.cuh"

namespace fb {

// (1-yf) * ((1-xf)*s00 + xf*s01) + yf * ((1-xf)*s10 + xf*s11), interpolation.c:899-900
__device__ __forceinline__ float bilinear_full(float wx0, float xf, float wy0, float yf, float s00, float s01, float s10, float s11)
{
    const float top = __fadd_rn(__fmul_rn(wx0, s00), __fmul_rn(xf, s01));
    const float bot = __fadd_rn(__fmul_rn(wx0, s10), __fmul_rn(xf, s11));
    return __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(yf, bot));
}

// u' = u*c - v*s ; v' = u*s + v*c in fp64, rounded to fp32 (interpolation.c:804-808)
__device__ __forceinline__ void rotate_uv(float& u, float& v, double c, double s)
{
    const double ud = (double)u, vd = (double)v;
    const double un = __dsub_rn(__dmul_rn(ud, c), __dmul_rn(vd, s));
    const double vn = __dadd_rn(__dmul_rn(ud, s), __dmul_rn(vd, c));
    u = __double2float_rn(un);
    v = __double2float_rn(vn);
}

// cubic convolution weights for a = -0.5: w[i] = sum_j T[j] * (M[j][i] / 2), accumulated from 0 in j order
// exactly as interpolation.c:962-968, 981-1000
__device__ __forceinline__ void cubic_weights(double t, double (&w)[4])
{
    const double M[4][4] = {{0., 1., 0., 0.}, {-.5, 0., .5, 0.}, {1., -2.5, 2., -.5}, {-.5, 1.5, -1.5, .5}};
    double T[4];
    T[0] = 1.;
    T[1] = t;
    T[2] = __dmul_rn(t, t);
    T[3] = __dmul_rn(T[2], t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double acc = 0.;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            acc = __dadd_rn(acc, __dmul_rn(T[j], M[j][i]));
        w[i] = acc;
    }
}

// One row of the 4x4 stencil: 0. + wx[0]*v0 + wx[1]*v1 + wx[2]*v2 + wx[3]*v3, left to right (interpolation.c:1006-1013).
// fma(w, v, +0.) IS the reference's "0. + w*v": the product is rounded once and adding zero is exact (a zero product
// gives +0 either way), so the first multiply-add costs one fp64 instruction instead of two.
__device__ __forceinline__ double bicubic_row(const double (&wx)[4], double v0, double v1, double v2, double v3)
{
    double row = __fma_rn(wx[0], v0, 0.);
    row = __dadd_rn(row, __dmul_rn(wx[1], v1));
    row = __dadd_rn(row, __dmul_rn(wx[2], v2));
    row = __dadd_rn(row, __dmul_rn(wx[3], v3));
    return row;
}

// outvalues[z] += XMF[i] * MY[i] with a float accumulator: re-rounded to fp32 after each of the four rows
// (interpolation.c:1005, :1017-1020).  FIRST: the accumulator is still +0.f.
template <bool FIRST>
__device__ __forceinline__ float bicubic_acc(float acc, double row, double wy)
{
    if (FIRST)
        return __double2float_rn(__fma_rn(row, wy, 0.));
    return __double2float_rn(__dadd_rn((double)acc, __dmul_rn(row, wy)));
}

// fp64 row sums, fp32 accumulator re-rounded after each of the four rows (interpolation.c:1002-1021); taps from
// global memory through the read-only path
__device__ __forceinline__ float bicubic_eval(const float* __restrict__ s, int ix, const double (&wx)[4], const double (&wy)[4])
{
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float* p = s + r * ix;
        const double row = bicubic_row(wx, (double)__ldg(p), (double)__ldg(p + 1), (double)__ldg(p + 2), (double)__ldg(p + 3));
        acc = (r == 0) ? bicubic_acc<true>(acc, row, wy[0]) : bicubic_acc<false>(acc, row, wy[r]);
    }
    return acc;
}

} // namespace fb
