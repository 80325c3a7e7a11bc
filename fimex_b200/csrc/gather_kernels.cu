// fimex_b200/csrc/gather_kernels.cu -- the per-slice hot path: K3 nearest neighbour, K4 bilinear,
// K5 bicubic gathers, with the K6 u/v rotation fused into the same pass (or stand-alone).
//
// Replaces CachedInterpolation::interpolateValues (/root/reference/src/CachedInterpolation.cc:118-147) and
// the three per-point functions it calls (src/interpolation.c:862-1028), plus
// mifi_vector_reproject_values_by_matrix_f (src/interpolation.c:790-812).
//
// Layout: input fp32 [z][iy][ix], output fp32 [z][oy][ox], x fastest (interpolation.h:423-426).  A thread
// owns VEC consecutive target points of one row-major output level and walks a chunk of levels, so that
//   * the compiled table entry of a point is read once per chunk (not once per level),
//   * the store of a warp is one contiguous 128*VEC-byte run per level (128-bit st.global.cs when VEC==4),
//   * source taps are read through the read-only path; neighbouring target points share source cells
//     (the target grid is normally finer than the source), so these hit L1/L2 and HBM sees each source
//     cell of the footprint once per level.
// Arithmetic is the reference's, operation for operation, with __fmul_rn/__fadd_rn (and the fp64
// equivalents) so that nvcc cannot contract to FMA: results are bit-identical to the reference built for
// x86-64, which has no FMA either (SURVEY.md 8a trap 7).
#include "kernels.h"
#include "interp_math.cuh"

#include <cstdlib>

namespace fb {

namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ float ldg_f(const float* p)
{
    return __ldg(p);
}

__device__ __forceinline__ float bilinear_any(int mode, const float* __restrict__ s, int ix, float wx0, float xf, float wy0, float yf)
{
    switch (mode) {
    case FB_BL_FULL:
        return bilinear_full(wx0, xf, wy0, yf, ldg_f(s), ldg_f(s + 1), ldg_f(s + ix), ldg_f(s + ix + 1));
    case FB_BL_XLIN: // :911
        return __fadd_rn(__fmul_rn(wx0, ldg_f(s)), __fmul_rn(xf, ldg_f(s + 1)));
    case FB_BL_YLIN: // :931
        return __fadd_rn(__fmul_rn(wy0, ldg_f(s)), __fmul_rn(yf, ldg_f(s + ix)));
    case FB_BL_NEAR: // :940
        return ldg_f(s);
    default:
        return undef_f();
    }
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* o, const float (&r)[VEC], long long valid)
{
    if constexpr (VEC == 4) {
        if (valid >= 4) {
            __stcs(reinterpret_cast<float4*>(o), make_float4(r[0], r[1], r[2], r[3]));
            return;
        }
    } else if constexpr (VEC == 2) {
        if (valid >= 2) {
            __stcs(reinterpret_cast<float2*>(o), make_float2(r[0], r[1]));
            return;
        }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k)
        if (k < valid)
            __stcs(o + k, r[k]);
}

struct ZRange {
    long long z0, z1;
};
__device__ __forceinline__ ZRange z_chunk(long long nz)
{
    const long long per = (nz + gridDim.y - 1) / gridDim.y;
    ZRange r;
    r.z0 = (long long)blockIdx.y * per;
    r.z1 = r.z0 + per < nz ? r.z0 + per : nz;
    return r;
}

// ------------------------------------------------------------------------------------------------ K3
// NFIELD = 1: scalar field; NFIELD = 2: u and v in one pass with optional rotation
template <int VEC, int NFIELD, bool ROT>
__global__ void __launch_bounds__(kThreads) k_gather_nn(GatherGeom g, const int* __restrict__ off_tab, const double2* __restrict__ cs,
                                                      const float* __restrict__ in0, const float* __restrict__ in1,
                                                      float* __restrict__ out0, float* __restrict__ out1)
{
    const long long q = (blockIdx.x * (long long)kThreads + threadIdx.x) * VEC;
    if (q >= g.out_level)
        return;
    const long long valid = g.out_level - q;
    int off[VEC];
    double2 rot[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        off[k] = (k < valid) ? __ldg(off_tab + q + k) : -1;
        if (ROT)
            rot[k] = (k < valid) ? __ldg(cs + q + k) : make_double2(1., 0.);
    }
    const ZRange zr = z_chunk(g.nz);
    const float* p0 = in0 + zr.z0 * g.in_level;
    const float* p1 = (NFIELD == 2) ? in1 + zr.z0 * g.in_level : nullptr;
    float* o0 = out0 + zr.z0 * g.out_level + q;
    float* o1 = (NFIELD == 2) ? out1 + zr.z0 * g.out_level + q : nullptr;
#pragma unroll 4
    for (long long z = zr.z0; z < zr.z1; ++z) {
        float a[VEC], b[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            a[k] = off[k] >= 0 ? ldg_f(p0 + off[k]) : undef_f();
            if (NFIELD == 2)
                b[k] = off[k] >= 0 ? ldg_f(p1 + off[k]) : undef_f();
        }
        if (ROT) {
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                rotate_uv(a[k], b[k], rot[k].x, rot[k].y);
        }
        store_vec<VEC>(o0, a, valid);
        p0 += g.in_level;
        o0 += g.out_level;
        if (NFIELD == 2) {
            store_vec<VEC>(o1, b, valid);
            p1 += g.in_level;
            o1 += g.out_level;
        }
    }
}

// ------------------------------------------------------------------------------------------------ K4
template <int VEC, int NFIELD, bool ROT>
__global__ void __launch_bounds__(kThreads) k_gather_bilinear(GatherGeom g, const int4* __restrict__ tab, const double2* __restrict__ cs,
                                                            const float* __restrict__ in0, const float* __restrict__ in1,
                                                            float* __restrict__ out0, float* __restrict__ out1)
{
    const long long q = (blockIdx.x * (long long)kThreads + threadIdx.x) * VEC;
    if (q >= g.out_level)
        return;
    const long long valid = g.out_level - q;
    int off[VEC], mode[VEC];
    float xf[VEC], yf[VEC], wx0[VEC], wy0[VEC];
    double2 rot[VEC];
    bool all_full = true;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        int4 e = make_int4(0, 0, 0, FB_BL_NAN);
        if (k < valid)
            e = __ldg(tab + q + k);
        off[k] = e.x;
        xf[k] = __int_as_float(e.y);
        yf[k] = __int_as_float(e.z);
        mode[k] = e.w;
        wx0[k] = __fsub_rn(1.f, xf[k]);
        wy0[k] = __fsub_rn(1.f, yf[k]);
        all_full = all_full && (e.w == FB_BL_FULL);
        if (ROT)
            rot[k] = (k < valid) ? __ldg(cs + q + k) : make_double2(1., 0.);
    }
    const ZRange zr = z_chunk(g.nz);
    const int ix = g.ix;
    const float* p0 = in0 + zr.z0 * g.in_level;
    const float* p1 = (NFIELD == 2) ? in1 + zr.z0 * g.in_level : nullptr;
    float* o0 = out0 + zr.z0 * g.out_level + q;
    float* o1 = (NFIELD == 2) ? out1 + zr.z0 * g.out_level + q : nullptr;
    if (all_full) {
#pragma unroll 2
        for (long long z = zr.z0; z < zr.z1; ++z) {
            float a[VEC], b[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const float* s = p0 + off[k];
                a[k] = bilinear_full(wx0[k], xf[k], wy0[k], yf[k], ldg_f(s), ldg_f(s + 1), ldg_f(s + ix), ldg_f(s + ix + 1));
                if (NFIELD == 2) {
                    const float* t = p1 + off[k];
                    b[k] = bilinear_full(wx0[k], xf[k], wy0[k], yf[k], ldg_f(t), ldg_f(t + 1), ldg_f(t + ix), ldg_f(t + ix + 1));
                }
            }
            if (ROT) {
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    rotate_uv(a[k], b[k], rot[k].x, rot[k].y);
            }
            store_vec<VEC>(o0, a, valid);
            p0 += g.in_level;
            o0 += g.out_level;
            if (NFIELD == 2) {
                store_vec<VEC>(o1, b, valid);
                p1 += g.in_level;
                o1 += g.out_level;
            }
        }
    } else { // grid edge: per-point mode
        for (long long z = zr.z0; z < zr.z1; ++z) {
            float a[VEC], b[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                a[k] = bilinear_any(mode[k], p0 + off[k], ix, wx0[k], xf[k], wy0[k], yf[k]);
                if (NFIELD == 2)
                    b[k] = bilinear_any(mode[k], p1 + off[k], ix, wx0[k], xf[k], wy0[k], yf[k]);
            }
            if (ROT) {
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    rotate_uv(a[k], b[k], rot[k].x, rot[k].y);
            }
            store_vec<VEC>(o0, a, valid);
            p0 += g.in_level;
            o0 += g.out_level;
            if (NFIELD == 2) {
                store_vec<VEC>(o1, b, valid);
                p1 += g.in_level;
                o1 += g.out_level;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ K5
template <int NFIELD, bool ROT>
__global__ void __launch_bounds__(kThreads) k_gather_bicubic(GatherGeom g, const int* __restrict__ off_tab, const double2* __restrict__ frac,
                                                           const double2* __restrict__ cs, const float* __restrict__ in0,
                                                           const float* __restrict__ in1, float* __restrict__ out0,
                                                           float* __restrict__ out1)
{
    const long long q = blockIdx.x * (long long)kThreads + threadIdx.x;
    if (q >= g.out_level)
        return;
    const int off = __ldg(off_tab + q);
    const ZRange zr = z_chunk(g.nz);
    float* o0 = out0 + zr.z0 * g.out_level + q;
    float* o1 = (NFIELD == 2) ? out1 + zr.z0 * g.out_level + q : nullptr;
    if (off < 0) { // outside the 4x4 support: NaN, no edge fallback (:1022-1026); rotating NaN stays NaN
        for (long long z = zr.z0; z < zr.z1; ++z) {
            __stcs(o0, undef_f());
            o0 += g.out_level;
            if (NFIELD == 2) {
                __stcs(o1, undef_f());
                o1 += g.out_level;
            }
        }
        return;
    }
    const double2 f = __ldg(frac + q);
    double wx[4], wy[4];
    cubic_weights(f.x, wx);
    cubic_weights(f.y, wy);
    double2 rot = make_double2(1., 0.);
    if (ROT)
        rot = __ldg(cs + q);
    const float* p0 = in0 + zr.z0 * g.in_level + off;
    const float* p1 = (NFIELD == 2) ? in1 + zr.z0 * g.in_level + off : nullptr;
    for (long long z = zr.z0; z < zr.z1; ++z) {
        float a = bicubic_eval(p0, g.ix, wx, wy);
        float b = 0.f;
        if (NFIELD == 2)
            b = bicubic_eval(p1, g.ix, wx, wy);
        if (ROT)
            rotate_uv(a, b, rot.x, rot.y);
        __stcs(o0, a);
        p0 += g.in_level;
        o0 += g.out_level;
        if (NFIELD == 2) {
            __stcs(o1, b);
            p1 += g.in_level;
            o1 += g.out_level;
        }
    }
}

// ------------------------------------------------------------------------------------------------ K6 stand-alone
__global__ void __launch_bounds__(kThreads) k_rotate(const double2* __restrict__ cs, float* __restrict__ u, float* __restrict__ v,
                                                   long long layer, long long nz)
{
    const long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
    if (i >= layer)
        return;
    const double2 m = __ldg(cs + i);
    const long long per = (nz + gridDim.y - 1) / gridDim.y;
    const long long z0 = (long long)blockIdx.y * per;
    const long long z1 = z0 + per < nz ? z0 + per : nz;
    float* pu = u + z0 * layer + i;
    float* pv = v + z0 * layer + i;
#pragma unroll 4
    for (long long z = z0; z < z1; ++z) {
        float a = *pu, b = *pv;
        rotate_uv(a, b, m.x, m.y);
        *pu = a;
        *pv = b;
        pu += layer;
        pv += layer;
    }
}

__global__ void __launch_bounds__(kThreads) k_rotate_direction(const double* __restrict__ matrix, float* __restrict__ angle,
                                                             long long layer, long long nz)
{
    const long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
    if (i >= layer)
        return;
    const double turn = __dmul_rn(FB_RAD_TO_DEG, matrix[4 * i + 3]);
    float* p = angle + i;
    for (long long z = 0; z < nz; ++z, p += layer) { // interpolation.c:823-832
        double a = __dsub_rn((double)*p, turn);
        if (a < 0)
            a = __dadd_rn(a, 360.);
        if (a > 360)
            a = __dsub_rn(a, 360.);
        *p = __double2float_rn(a);
    }
}

bool aligned16(const void* p)
{
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

} // namespace

// Number of level chunks (gridDim.y).  A CTA walks `chunk` consecutive levels of its target points, so the table
// entry of a point is read once per chunk.  Measured on B200 (profiles/): chunks of 32..96 levels run at the same
// speed, longer ones are slower (CTAs drift apart in z and the set of DRAM pages being written at any moment
// grows: 658-level chunks cost 40 % more time per level than 64-level chunks), shorter ones re-read the tables too
// often.  So: ~64 levels per chunk (`levels`; the staged bilinear gather is 1.5 % faster with 128), and at least ~4 CTAs per
// SM in flight for small slices.
int z_chunks(long long ctas_x, long long nz, int levels)
{
    long long gy = (nz + levels - 1) / levels;
    const long long min_ctas = (long long)sm_count() * 4;
    if (ctas_x * gy < min_ctas) {
        const long long more = (min_ctas + ctas_x - 1) / ctas_x;
        const long long cap = nz / 8 > 1 ? nz / 8 : 1; // never below 8 levels per chunk
        gy = more < cap ? more : cap;
    }
    if (const char* env = std::getenv("FIMEX_B200_ZCHUNK")) { // experiments only: levels per CTA
        const long long want_chunk = std::atoll(env);
        if (want_chunk > 0)
            gy = (nz + want_chunk - 1) / want_chunk;
    }
    if (gy < 1)
        gy = 1;
    if (gy > 65535)
        gy = 65535;
    return (int)gy;
}

// ------------------------------------------------------------------------------------------------- launchers
#define FB_GATHER_PRECHECK()                                                                                                              \
    if (g.out_level == 0 || g.nz == 0)                                                                                                     \
        return FB_OK;                                                                                                                      \
    FB_REQUIRE(g.in_level < 2147483647LL, "source level larger than 2^31 cells")

int launch_gather_nn(const GatherGeom& g, const int* d_off, const float* d_in, float* d_out, cudaStream_t st)
{
    FB_GATHER_PRECHECK();
    if ((g.out_level % 4) == 0 && aligned16(d_out)) {
        const int gx = ceil_div(ceil_div(g.out_level, 4), kThreads);
        dim3 grid(gx, z_chunks(gx, g.nz));
        k_gather_nn<4, 1, false><<<grid, kThreads, 0, st>>>(g, d_off, nullptr, d_in, nullptr, d_out, nullptr);
    } else {
        const int gx = ceil_div(g.out_level, kThreads);
        dim3 grid(gx, z_chunks(gx, g.nz));
        k_gather_nn<1, 1, false><<<grid, kThreads, 0, st>>>(g, d_off, nullptr, d_in, nullptr, d_out, nullptr);
    }
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_gather_bilinear(const GatherGeom& g, const int4* d_tab, const float* d_in, float* d_out, cudaStream_t st)
{
    FB_GATHER_PRECHECK();
    if ((g.out_level % 4) == 0 && aligned16(d_out)) {
        const int gx = ceil_div(ceil_div(g.out_level, 4), kThreads);
        dim3 grid(gx, z_chunks(gx, g.nz));
        k_gather_bilinear<4, 1, false><<<grid, kThreads, 0, st>>>(g, d_tab, nullptr, d_in, nullptr, d_out, nullptr);
    } else {
        const int gx = ceil_div(g.out_level, kThreads);
        dim3 grid(gx, z_chunks(gx, g.nz));
        k_gather_bilinear<1, 1, false><<<grid, kThreads, 0, st>>>(g, d_tab, nullptr, d_in, nullptr, d_out, nullptr);
    }
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_gather_bicubic(const GatherGeom& g, const int* d_off, const double2* d_frac, const float* d_in, float* d_out, cudaStream_t st)
{
    FB_GATHER_PRECHECK();
    const int gx = ceil_div(g.out_level, kThreads);
    dim3 grid(gx, z_chunks(gx, g.nz));
    k_gather_bicubic<1, false><<<grid, kThreads, 0, st>>>(g, d_off, d_frac, nullptr, d_in, nullptr, d_out, nullptr);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_gather_vector(int method, const GatherGeom& g, const void* d_tab, const void* d_tab2, const double2* d_cs, const float* d_u_in,
                         const float* d_v_in, float* d_u_out, float* d_v_out, cudaStream_t st)
{
    FB_GATHER_PRECHECK();
    const bool rot = d_cs != nullptr;
    const bool vec2 = (g.out_level % 2) == 0 && ((reinterpret_cast<uintptr_t>(d_u_out) | reinterpret_cast<uintptr_t>(d_v_out)) & 7u) == 0;
    if (method == FB_BICUBIC) {
        const int gx = ceil_div(g.out_level, kThreads);
        dim3 grid(gx, z_chunks(gx, g.nz));
        if (rot)
            k_gather_bicubic<2, true><<<grid, kThreads, 0, st>>>(g, (const int*)d_tab, (const double2*)d_tab2, d_cs, d_u_in, d_v_in, d_u_out,
                                                                 d_v_out);
        else
            k_gather_bicubic<2, false><<<grid, kThreads, 0, st>>>(g, (const int*)d_tab, (const double2*)d_tab2, nullptr, d_u_in, d_v_in,
                                                                  d_u_out, d_v_out);
    } else if (method == FB_BILINEAR) {
        const int vec = vec2 ? 2 : 1;
        const int gx = ceil_div(ceil_div(g.out_level, vec), kThreads);
        dim3 grid(gx, z_chunks(gx, g.nz));
        const int4* t = (const int4*)d_tab;
        if (vec2 && rot)
            k_gather_bilinear<2, 2, true><<<grid, kThreads, 0, st>>>(g, t, d_cs, d_u_in, d_v_in, d_u_out, d_v_out);
        else if (vec2)
            k_gather_bilinear<2, 2, false><<<grid, kThreads, 0, st>>>(g, t, nullptr, d_u_in, d_v_in, d_u_out, d_v_out);
        else if (rot)
            k_gather_bilinear<1, 2, true><<<grid, kThreads, 0, st>>>(g, t, d_cs, d_u_in, d_v_in, d_u_out, d_v_out);
        else
            k_gather_bilinear<1, 2, false><<<grid, kThreads, 0, st>>>(g, t, nullptr, d_u_in, d_v_in, d_u_out, d_v_out);
    } else { // nearest neighbour family
        const int vec = vec2 ? 2 : 1;
        const int gx = ceil_div(ceil_div(g.out_level, vec), kThreads);
        dim3 grid(gx, z_chunks(gx, g.nz));
        const int* t = (const int*)d_tab;
        if (vec2 && rot)
            k_gather_nn<2, 2, true><<<grid, kThreads, 0, st>>>(g, t, d_cs, d_u_in, d_v_in, d_u_out, d_v_out);
        else if (vec2)
            k_gather_nn<2, 2, false><<<grid, kThreads, 0, st>>>(g, t, nullptr, d_u_in, d_v_in, d_u_out, d_v_out);
        else if (rot)
            k_gather_nn<1, 2, true><<<grid, kThreads, 0, st>>>(g, t, d_cs, d_u_in, d_v_in, d_u_out, d_v_out);
        else
            k_gather_nn<1, 2, false><<<grid, kThreads, 0, st>>>(g, t, nullptr, d_u_in, d_v_in, d_u_out, d_v_out);
    }
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_rotate(const double2* d_cs, float* d_u, float* d_v, long long layer, long long nz, cudaStream_t st)
{
    if (layer == 0 || nz == 0)
        return FB_OK;
    const int gx = ceil_div(layer, kThreads);
    dim3 grid(gx, z_chunks(gx, nz));
    k_rotate<<<grid, kThreads, 0, st>>>(d_cs, d_u, d_v, layer, nz);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_rotate_direction(const double* d_matrix, float* d_angle, long long layer, long long nz, cudaStream_t st)
{
    if (layer == 0 || nz == 0)
        return FB_OK;
    k_rotate_direction<<<ceil_div(layer, kThreads), kThreads, 0, st>>>(d_matrix, d_angle, layer, nz);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

} // namespace fb
