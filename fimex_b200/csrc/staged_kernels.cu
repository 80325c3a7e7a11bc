// fimex_b200/csrc/staged_kernels.cu -- K4 (bilinear) as a shared-memory staged gather: the fast path of
// CachedInterpolation::interpolateValues (/root/reference/src/CachedInterpolation.cc:118-147 with
// mifi_get_values_bilinear_f, src/interpolation.c:881-957).
//
// Why: the direct gather (gather_kernels.cu) issues four 4-byte global loads per output value and is bound by
// L1/LSU request rate and load latency (ncu, profiles/): ~33 % of the HBM roofline.  The target grid is
// normally finer than the source, so the 1024 target points of a tile touch only a few dozen distinct source
// values.  Here a CTA owns a tile of 1024 target points (64 x 16 for bilinear, 128 x 8 for nearest neighbour) and, per batch of
// up to 8 levels,
//   1. copies the tile's DISTINCT source values ("taps", a sorted list of offsets inside a level) into shared
//      memory with cp.async (LDGSTS), double buffered: batch b+1 is in flight while batch b is consumed, so
//      each source value is requested from L2/HBM once per tile and level, not once per target point.  The staging
//      buffer is tap-major (tap r, level zi at r*12 + zi), so the levels of a tap are contiguous;
//   2. every thread then reads four levels of each of its taps with one 128-bit shared load (two row indices per
//      point; the right-hand neighbour is the next list entry), evaluates the reference's formula (no FMA
//      contraction: bit-identical to the CPU) and streams its outputs with st.global.cs: bilinear lane = x, a warp
//      writing 128 contiguous bytes of a row; nearest neighbour 4 consecutive points per thread with one 128-bit store,
//      a warp writing 512 contiguous bytes.
// The tap list is per tile, not a bounding box, so tiles over the pole or across the 0/360 longitude seam
// (where the footprint of a tile is scattered) cost no more than their number of distinct taps.
//
// Tile table (built once per grid by k_compile_tiles, on the device):
//   taps  [tile][4096]  int    sorted distinct source offsets y*ix + x
//   ntaps [tile]        int
//   meta  [tile][256]   uint4  per thread, per point: a | b << 12 | mode << 24, a/b = list index of the tap at
//                              (x0, y0) / (x0, y0+1) of that point
//   xf/yf [tile][256]   float4 per thread: the reference's float xfrac / yfrac of its 4 points
#include "kernels.h"
#include "tables.cuh"
#include "convert.cuh"
#include "interp_math.cuh"

#include <cuda.h> // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)

#include <cstdlib>
#include <cstring>
#include <vector>

namespace fb {

namespace {

#ifndef FB_CTA_THREADS
#define FB_CTA_THREADS 256 // threads per CTA = quarter of the tile's points (experiments: -DFB_CTA_THREADS=128)
#endif
constexpr int kThreads = FB_CTA_THREADS;
constexpr int kTilePts = 4 * kThreads;
// Tile shapes (1024 target points, 4 per thread).  Measured on B200 (profiles/ubench/store_bw.cu, pure stores, 64-level
// chunks): tiles whose rows are 128 B wide reach 5.66 TB/s, 256 B rows 6.35 TB/s, 512 B rows written with 128-bit stores
// 7.1 TB/s -- the wider the contiguous run a CTA writes per row, the fewer DRAM pages are open at once.  The width a
// bilinear warp can cover is limited by shared-memory banks instead: the 32 lanes of a load should see fewer than 32
// distinct taps, i.e. span fewer than ~30 source cells.
//   bilinear: 64 x 16, lane = x, thread t owns (t & 63, (t >> 6) + 4k): a row is written by two neighbouring warps.  (Giving a
//   thread two x-neighbours in each of two rows would let a warp write 256-byte row segments with 64-bit stores -- 7.1 TB/s in
//   the store benchmark instead of 6.4 -- but its shared loads then see ~26 distinct taps per warp instead of ~13 and the
//   kernel is 17 % slower: 13.1 ms against 11.2.)
//   nearest neighbour: 128 x 8, thread t owns (4 (t & 31) + k, t >> 5): one 128-bit store per thread and level
#ifndef FB_STAGE_BUFFERS
#define FB_STAGE_BUFFERS 2 // staging buffers of the cp.async pipeline (experiments: -DFB_STAGE_BUFFERS=3)
#endif
constexpr int kBuffers = FB_STAGE_BUFFERS;
#ifndef FB_NN_CTAS
#define FB_NN_CTAS 3 // resident CTAs per SM the nearest-neighbour kernel is compiled for (experiments: -DFB_NN_CTAS=4)
#endif
#ifndef FB_BLQ_CTAS
#define FB_BLQ_CTAS 3 // resident CTAs per SM the quad-layout bilinear kernel is compiled for (experiments: -DFB_BLQ_CTAS=2)
#endif
#ifndef FB_BLQ_LV
#define FB_BLQ_LV 4 // levels per shared load of the quad-layout bilinear kernel: 4 (LDS.128) or 2 (LDS.64, 16 fewer live registers)
#endif
#ifndef FB_BL_TILE_X
#define FB_BL_TILE_X 64 // bilinear tile width: 32 or 64 (experiments: -DFB_BL_TILE_X=32)
#endif
// Tile<QUAD>: QUAD = a thread owns 4 x-neighbours of one row (128 x 8 tiles: nearest neighbour; bilinear with
// FIMEX_B200_BILINEAR_QUAD=1); !QUAD = lane = x, a thread owns one column of 4 rows (64 x 16 tiles: bilinear by default)
template <bool QUAD>
struct Tile {
    static constexpr int X = QUAD ? 128 : FB_BL_TILE_X, Y = kTilePts / X;
    static constexpr int RowStep = QUAD ? 0 : kThreads / X; // column layout: rows between point k and point k+1 of a thread
    // position of point k of thread t inside the tile
    static __device__ __forceinline__ int px(int t, int k) { return QUAD ? 4 * (t & 31) + k : (t % X); }
    static __device__ __forceinline__ int py(int t, int k) { return QUAD ? (t >> 5) : (t / X) + RowStep * k; }
};
constexpr int kMaxTaps = 4 * kTilePts;                              // worst case: every point has its own 4 taps
constexpr int kStageFloats = 16 * kThreads;                          // one staging buffer (16 KB), two of them
constexpr int kMaxBatch = 8;                                        // levels staged per barrier
constexpr int kNoTap = 0x7fffffff;
// tap-major staging: element (tap r, level zi) at r*kLvlStride + zi.  12 words keep every tap 16-byte aligned, and the taps of
// the 8 lanes of a quarter warp (consecutive r) start in banks 0, 12, 24, 4, 16, 28, 8, 20: four banks each, no conflict.
constexpr int kLvlStride = 12;
constexpr int kFastTaps = kStageFloats / kLvlStride;                // tiles with at most 341 taps use it (all but pole/seam tiles)
#ifndef FB_NN_BULK_DEFAULT
#define FB_NN_BULK_DEFAULT 0 // nearest neighbour: 0 = per-thread stores, 1 / 2 = output tile stored by the copy engine (FIMEX_B200_NN_BULK)
#endif
constexpr int kNnBulkDefault = FB_NN_BULK_DEFAULT;

// ------------------------------------------------------------------------------------------------ table compiler
template <bool NN, bool QUAD>
__global__ void __launch_bounds__(kThreads) k_compile_tiles(const double* __restrict__ px, const double* __restrict__ py, int ox, int oy,
                                                          int ix, int iy, int tiles_x, int* __restrict__ taps, int* __restrict__ ntaps,
                                                          uint4* __restrict__ meta, float4* __restrict__ xf4, float4* __restrict__ yf4)
{
    __shared__ int s_keys[kMaxTaps];
    __shared__ int s_uniq[kMaxTaps];
    __shared__ int s_warp_tot[kThreads / 32];
    const int tile = blockIdx.x;
    const int t = threadIdx.x;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    int off[4], mode[4];
    float xf[4], yf[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int4 e = make_int4(0, 0, 0, FB_BL_NAN);
        const int x = tx * Tile<QUAD>::X + Tile<QUAD>::px(t, k), y = ty * Tile<QUAD>::Y + Tile<QUAD>::py(t, k);
        if (y < oy && x < ox) {
            const long long i = (long long)y * ox + x;
            e = NN ? classify_nn(px[i], py[i], ix, iy) : classify_bilinear(px[i], py[i], ix, iy);
        }
        off[k] = e.x;
        xf[k] = __int_as_float(e.y);
        yf[k] = __int_as_float(e.z);
        mode[k] = e.w;
        // the taps this point reads (interpolation.c:894-897, :909-910, :929-930, :940)
        const bool any = e.w != FB_BL_NAN;
        const bool right = e.w == FB_BL_FULL || e.w == FB_BL_XLIN;
        const bool down = e.w == FB_BL_FULL || e.w == FB_BL_YLIN;
        int* kk = s_keys + (t * 4 + k) * 4;
        kk[0] = any ? e.x : kNoTap;
        kk[1] = right ? e.x + 1 : kNoTap;
        kk[2] = down ? e.x + ix : kNoTap;
        kk[3] = (right && down) ? e.x + ix + 1 : kNoTap;
    }
    __syncthreads();
    // bitonic sort of the 4096 keys, ascending
    for (int k = 2; k <= kMaxTaps; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < kMaxTaps; i += kThreads) {
                const int p = i ^ j;
                if (p > i) {
                    const int a = s_keys[i], b = s_keys[p];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        s_keys[i] = b;
                        s_keys[p] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    // distinct keys -> s_uniq (block-wide exclusive scan of the per-thread counts; 16 keys per thread)
    unsigned flags = 0;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int i = t * 16 + k;
        const int v = s_keys[i];
        const bool first = (v != kNoTap) && (i == 0 || s_keys[i - 1] != v);
        flags |= first ? (1u << k) : 0u;
        cnt += first;
    }
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= o)
            incl += n;
    }
    if ((t & 31) == 31)
        s_warp_tot[t >> 5] = incl;
    __syncthreads();
    int base = 0, total = 0;
    for (int w = 0; w < kThreads / 32; ++w) {
        if (w < (t >> 5))
            base += s_warp_tot[w];
        total += s_warp_tot[w];
    }
    int pos = base + incl - cnt;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (flags & (1u << k))
            s_uniq[pos++] = s_keys[t * 16 + k];
    __syncthreads();
    for (int j = t; j < total; j += kThreads)
        taps[(size_t)tile * kMaxTaps + j] = s_uniq[j];
    if (t == 0)
        ntaps[tile] = total;
    // each point looks up its two row starts in the sorted list (the right-hand taps are the next entries)
    auto find = [&](int key) {
        int lo = 0, hi = total - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_uniq[mid] < key)
                lo = mid + 1;
            else
                hi = mid;
        }
        return lo;
    };
    unsigned m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned a = 0, b = 0;
        if (mode[k] != FB_BL_NAN) {
            a = (unsigned)find(off[k]);
            b = (mode[k] == FB_BL_FULL || mode[k] == FB_BL_YLIN) ? (unsigned)find(off[k] + ix) : a;
        }
        m[k] = a | (b << 12) | ((unsigned)mode[k] << 24);
    }
    const size_t slot = (size_t)tile * kThreads + t;
    meta[slot] = make_uint4(m[0], m[1], m[2], m[3]);
    if (!NN) {
        xf4[slot] = make_float4(xf[0], xf[1], xf[2], xf[3]);
        yf4[slot] = make_float4(yf[0], yf[1], yf[2], yf[3]);
    }
}

// ------------------------------------------------------------------------------------------------ gather
// 4-byte asynchronous global->shared copy (LDGSTS)
__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gmem_src)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_pending()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Out: StorePlain (interpolateValues' own float output) or StoreAs<T> (interpolationArray2Data fused into the store:
// NaN -> fill, round + cast to the variable's type).  fill_in: mifi_bad2nanf fused into the staging (values equal to
// bad0 / bad1 become NaN before any point reads them).  NF = 2: both components of a vector through one table pass
// (CDMInterpolator.cc:255-276), ROT: rotated in the epilogue (mifi_vector_reproject_values_by_matrix_f, interpolation.c:790-812).
// QUAD: geometry (see Tile).  NN implies QUAD.  For bilinear, QUAD also switches the common branch to the tap-reuse form:
// the 4 points of a thread are x-neighbours, so point k mostly reads the cell of point k-1 again (nothing is loaded) or the
// cell one column to the right (its left taps are the right taps already in registers: two loads instead of four).
template <bool NN, int NF, bool ROT, class Out, bool QUAD>
__global__ void __launch_bounds__(kThreads, (NF == 1 ? (NN ? FB_NN_CTAS : (QUAD ? FB_BLQ_CTAS : 3)) : 2) * (256 / kThreads))
    k_gather_bilinear_staged(GatherGeom g, int tiles_x, const int* __restrict__ taps, const int* __restrict__ ntaps_tab,
                             const uint4* __restrict__ meta, const float4* __restrict__ xf4, const float4* __restrict__ yf4,
                             const float* __restrict__ in0, const float* __restrict__ in1, typename Out::type* __restrict__ out0,
                             typename Out::type* __restrict__ out1, const double2* __restrict__ cs, Out conv, int fill_in, float bad0,
                             float bad1, int vec_ok, const int* __restrict__ tile_list)
{
    extern __shared__ __align__(16) float s_dyn[]; // [kBuffers][NF fields][kStageFloats]: later batches land while batch b is consumed
    auto stage = [&](int buf, int f) { return s_dyn + (buf * NF + f) * kStageFloats; };
    // tile_list: this launch covers only the listed tiles (the many-tap tiles the bulk-store kernel leaves out)
    const int tile = tile_list ? __ldg(tile_list + blockIdx.x) : (int)blockIdx.x;
    const int t = threadIdx.x;
    const int ntaps = __ldg(ntaps_tab + tile);
    const int* my_taps = taps + (size_t)tile * kMaxTaps;
    const size_t slot = (size_t)tile * kThreads + t;
    const uint4 m = __ldg(meta + slot);
    const float4 fx = NN ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(xf4 + slot);
    const float4 fy = NN ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(yf4 + slot);
    const unsigned mm[4] = {m.x, m.y, m.z, m.w};
    const float xf[4] = {fx.x, fx.y, fx.z, fx.w}, yf[4] = {fy.x, fy.y, fy.z, fy.w};
    int ia[4], ib[4], mode[4];
    float wx0[4], wy0[4];
    bool all_full = true; // every point of this thread takes the common branch: 4 taps (bilinear) / 1 tap (nearest neighbour)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ia[k] = (int)(mm[k] & 0xfffu);
        ib[k] = (int)((mm[k] >> 12) & 0xfffu);
        mode[k] = (int)(mm[k] >> 24);
        wx0[k] = __fsub_rn(1.f, xf[k]);
        wy0[k] = __fsub_rn(1.f, yf[k]);
        all_full = all_full && (mode[k] == (NN ? FB_BL_NEAR : FB_BL_FULL));
    }
    // bilinear in the quad layout: how point k's cell relates to point k-1's (list indices: the right-hand neighbour of a tap is
    // the next list entry).  same: nothing to load; shift: one column to the right, the old right taps become the left taps
    bool same[4] = {false, false, false, false}, shift[4] = {false, false, false, false};
    if (QUAD && !NN) {
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            same[k] = ia[k] == ia[k - 1] && ib[k] == ib[k - 1];
            shift[k] = ia[k] == ia[k - 1] + 1 && ib[k] == ib[k - 1] + 1;
        }
    }
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    // which of the thread's 4 points exist (the grid need not be a multiple of the tile), and their element offsets in a level
    static_assert(QUAD || !NN, "nearest neighbour uses the quad geometry");
    const int x0 = tx * Tile<QUAD>::X + Tile<QUAD>::px(t, 0), y0 = ty * Tile<QUAD>::Y + Tile<QUAD>::py(t, 0);
    const unsigned rstep = (unsigned)Tile<QUAD>::RowStep * (unsigned)g.ox; // column layout: from point k to point k+1 of a thread
    auto poff = [&](int k) -> unsigned { return QUAD ? (unsigned)k : (unsigned)k * rstep; };
    // points k < nvalid exist (k runs along x in the quad layout, down the rows in the column layout)
    int nvalid = 0;
    if (y0 < g.oy && x0 < g.ox)
        nvalid = QUAD ? g.ox - x0 : (g.oy - y0 + Tile<QUAD>::RowStep - 1) / (QUAD ? 1 : Tile<QUAD>::RowStep);
    const unsigned vmask = nvalid >= 4 ? 0xfu : (1u << nvalid) - 1u;
    const long long per = (g.nz + gridDim.y - 1) / gridDim.y;
    const long long z0 = (long long)blockIdx.y * per;
    const long long z1 = z0 + per < g.nz ? z0 + per : g.nz;
    const unsigned off0 = (unsigned)y0 * (unsigned)g.ox + (unsigned)x0; // point 0
    double2 rot[4];
    if (ROT) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            rot[k] = ((vmask >> k) & 1u) ? __ldg(cs + off0 + poff(k)) : make_double2(1., 0.);
    }
    // Two staging layouts.  fast: tap-major, element (tap r, level zi) at r*12 + zi -- a thread's eight row pointers are then
    // constant for a whole batch and the level is an immediate offset of the shared load (no address arithmetic in the
    // inner loop).  Tiles with more than 341 taps (over the pole, across the seam) keep the level-major layout with as many
    // levels per batch as fit.
    const bool fast = ntaps <= kFastTaps;
    const int zb = fast ? kMaxBatch : (kStageFloats / ntaps);
    // the first staging element of this thread is the same for every level; larger tiles loop over the rest
    const int tap0 = (t < ntaps) ? __ldg(my_taps + t) : -1;

    auto issue = [&](int buf, long long z, int nb) {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            float* dst = stage(buf, f);
            const float* lv = (f == 0 ? in0 : in1) + z * g.in_level;
            if (fast) {
                if (tap0 >= 0) { // warps past the end of the tap list (most of them: ~50 taps per tile) skip the loop entirely
                    const float* src = lv + tap0;
                    float* d = dst + t * kLvlStride;
#pragma unroll
                    for (int zi = 0; zi < kMaxBatch; ++zi)
                        if (zi < nb)
                            cp_async_f32(d + zi, src + zi * g.in_level);
                }
                if (ntaps > kThreads) {
                    for (int r = t + kThreads; r < ntaps; r += kThreads) {
                        const float* src = lv + __ldg(my_taps + r);
                        for (int zi = 0; zi < nb; ++zi)
                            cp_async_f32(dst + r * kLvlStride + zi, src + zi * g.in_level);
                    }
                }
            } else {
                for (int zi = 0; zi < nb; ++zi, lv += g.in_level, dst += ntaps) {
                    if (tap0 >= 0)
                        cp_async_f32(dst + t, lv + tap0);
                    for (int r = t + kThreads; r < ntaps; r += kThreads)
                        cp_async_f32(dst + r, lv + __ldg(my_taps + r));
                }
            }
        }
        cp_async_commit();
    };

    // mifi_bad2nanf on the batch that has just landed: every thread patches the elements it copied itself (its own
    // cp.async writes are visible to it after the wait), so no extra barrier is needed
    auto patch = [&](int buf, int nb) {
        const float nanv = undef_f();
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            float* dst = stage(buf, f);
            const float bad = f == 0 ? bad0 : bad1;
            if (fast) {
                for (int r = t; r < ntaps; r += kThreads) {
                    float* d = dst + r * kLvlStride;
                    for (int zi = 0; zi < nb; ++zi)
                        if (d[zi] == bad)
                            d[zi] = nanv;
                }
            } else {
                for (int zi = 0; zi < nb; ++zi, dst += ntaps)
                    for (int r = t; r < ntaps; r += kThreads)
                        if (dst[r] == bad)
                            dst[r] = nanv;
            }
        }
    };

    // kBuffers-deep pipeline: batches b+1 .. b+kBuffers-1 are in flight while batch b is consumed.  Every iteration commits
    // exactly one cp.async group (possibly empty), so "all but the newest kBuffers-2 groups have landed" is the wait condition.
    auto issue_or_skip = [&](int buf, long long z) {
        if (z < z1)
            issue(buf, z, (int)((z1 - z) < zb ? (z1 - z) : zb));
        else
            cp_async_commit();
    };
#pragma unroll
    for (int b = 0; b < kBuffers - 1; ++b)
        issue_or_skip(b, z0 + (long long)b * zb);
    int buf = 0;
    for (long long z = z0; z < z1; z += zb, buf = (buf + 1 == kBuffers ? 0 : buf + 1)) {
        const int nb = (int)((z1 - z) < zb ? (z1 - z) : zb);
        cp_async_wait_pending<kBuffers - 2>();
        if (fill_in)
            patch(buf, nb);
        __syncthreads(); // batch z has landed for every thread, and every thread is done reading the buffer refilled next
        issue_or_skip(buf == 0 ? kBuffers - 1 : buf - 1, z + (long long)(kBuffers - 1) * zb);
        if (vmask == 0)
            continue;
        const float* lvl = stage(buf, 0);
        typename Out::type* base0 = out0 + z * g.out_level; // uniform across the CTA
        typename Out::type* base1 = NF == 2 ? out1 + z * g.out_level : nullptr;
        if (fast && all_full && vmask == 0xfu) {
            const float* pa[4];
            const float* pb[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                pa[k] = lvl + ia[k] * kLvlStride;
                pb[k] = lvl + ib[k] * kLvlStride;
            }
            // Several levels at a time: the staged levels of a tap are contiguous, so one LDS.128 brings four of them -- a quarter
            // of the load instructions (9 % fewer instructions overall).  The L1TEX cycles stay what they were: 4.45 shared-load
            // wavefronts per 32 outputs and level either way (ncu), i.e. the merging of equal addresses that a 128-bit load
            // shows in profiles/ubench/lds_width.cu does not happen for the irregular runs of equal taps a rotated grid produces.
            // LV levels per step: 4; 2 for a rotated vector pair, whose 2 x 4 x LV results and the fp64 rotation
            // have to fit the 128 registers of 2 CTAs per SM
            constexpr int LV = (NF == 2 && ROT) ? 2 : ((QUAD && !NN && NF == 1) ? FB_BLQ_LV : 4);
            auto ldv = [](const float* p, float (&v)[LV]) {
                if constexpr (LV == 4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(p);
                    v[0] = t4.x, v[1] = t4.y, v[2] = t4.z, v[3] = t4.w;
                } else {
                    const float2 t2 = *reinterpret_cast<const float2*>(p);
                    v[0] = t2.x, v[1] = t2.y;
                }
            };
            auto step = [&](int q, int nlev) { // levels LV*q .. LV*q + LV-1 of the batch, the first nlev of them are stored
                float r[NF][4][LV]; // [field][point][level]
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    if (QUAD && !NN) { // x-neighbours: keep the taps of the previous point and load only what changed
                        float a0[LV], a1[LV], b0[LV], b1[LV];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float* qa = pa[k] + f * kStageFloats + LV * q;
                            const float* qb = pb[k] + f * kStageFloats + LV * q;
                            if (k > 0 && shift[k]) {
#pragma unroll
                                for (int l = 0; l < LV; ++l)
                                    a0[l] = a1[l], b0[l] = b1[l];
                            }
                            if (k == 0 || !same[k]) { // right-hand taps: a new cell, whichever way it was reached
                                ldv(qa + kLvlStride, a1);
                                ldv(qb + kLvlStride, b1);
                            }
                            if (k == 0 || !(same[k] || shift[k])) { // left-hand taps: only when the cell is not the neighbour's
                                ldv(qa, a0);
                                ldv(qb, b0);
                            }
#pragma unroll
                            for (int l = 0; l < LV; ++l)
                                r[f][k][l] = bilinear_full(wx0[k], xf[k], wy0[k], yf[k], a0[l], a1[l], b0[l], b1[l]);
                        }
                        continue;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float* qa = pa[k] + f * kStageFloats + LV * q;
                        const float* qb = pb[k] + f * kStageFloats + LV * q;
                        if (NN) { // copied value, bit for bit (interpolation.c:869-871)
                            ldv(qa, r[f][k]);
                        } else {
                            float a0[LV], a1[LV], b0[LV], b1[LV];
                            ldv(qa, a0), ldv(qa + kLvlStride, a1), ldv(qb, b0), ldv(qb + kLvlStride, b1);
#pragma unroll
                            for (int l = 0; l < LV; ++l)
                                r[f][k][l] = bilinear_full(wx0[k], xf[k], wy0[k], yf[k], a0[l], a1[l], b0[l], b1[l]);
                        }
                    }
                }
                if (ROT) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int l = 0; l < LV; ++l)
                            rotate_uv(r[0][k][l], r[NF - 1][k][l], rot[k].x, rot[k].y);
                }
#pragma unroll
                for (int l = 0; l < LV; ++l) {
                    if (l < nlev) {
#pragma unroll
                        for (int f = 0; f < NF; ++f) {
                            typename Out::type* dst = (f == 0 ? base0 : base1) + off0;
                            if (QUAD && vec_ok) {
                                store_vec4<typename Out::type>(dst, conv(r[f][0][l]), conv(r[f][1][l]), conv(r[f][2][l]), conv(r[f][3][l]));
                            } else {
                                __stcs(dst, conv(r[f][0][l]));
                                __stcs(dst + poff(1), conv(r[f][1][l]));
                                __stcs(dst + poff(2), conv(r[f][2][l]));
                                __stcs(dst + poff(3), conv(r[f][3][l]));
                            }
                        }
                        base0 += g.out_level;
                        if (NF == 2)
                            base1 += g.out_level;
                    }
                }
            };
            static_assert(kMaxBatch % LV == 0, "whole steps per batch");
            // (rotated vector pairs: a real loop -- unrolled, the scheduler hoists the loads of later steps and spills)
            if (nb == kMaxBatch) { // full batch: no per-level test
#pragma unroll((NF == 2 && ROT) ? 1 : kMaxBatch / LV)
                for (int q = 0; q < kMaxBatch / LV; ++q)
                    step(q, LV);
            } else { // levels past nb hold stale values of an earlier batch: computed, never stored
#pragma unroll((NF == 2 && ROT) ? 1 : kMaxBatch / LV)
                for (int q = 0; q < kMaxBatch / LV; ++q)
                    if (LV * q < nb)
                        step(q, nb - LV * q);
            }
        } else { // grid edge, partial tile or a tile with many taps: per-point mode (interpolation.c:904-953)
            const int sr = fast ? kLvlStride : 1;      // stride between neighbouring taps
            const int sz = fast ? 1 : ntaps;           // stride between levels
            for (int zi = 0; zi < nb; ++zi, lvl += sz) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float v[NF];
#pragma unroll
                    for (int f = 0; f < NF; ++f) {
                        v[f] = undef_f();
                        const float* qa = lvl + f * kStageFloats + ia[k] * sr;
                        const float* qb = lvl + f * kStageFloats + ib[k] * sr;
                        switch (mode[k]) {
                        case FB_BL_FULL: {
                            const float top = __fadd_rn(__fmul_rn(wx0[k], qa[0]), __fmul_rn(xf[k], qa[sr]));
                            const float bot = __fadd_rn(__fmul_rn(wx0[k], qb[0]), __fmul_rn(xf[k], qb[sr]));
                            v[f] = __fadd_rn(__fmul_rn(wy0[k], top), __fmul_rn(yf[k], bot));
                            break;
                        }
                        case FB_BL_XLIN:
                            v[f] = __fadd_rn(__fmul_rn(wx0[k], qa[0]), __fmul_rn(xf[k], qa[sr]));
                            break;
                        case FB_BL_YLIN:
                            v[f] = __fadd_rn(__fmul_rn(wy0[k], qa[0]), __fmul_rn(yf[k], qb[0]));
                            break;
                        case FB_BL_NEAR:
                            v[f] = qa[0];
                            break;
                        default:
                            break;
                        }
                    }
                    if (ROT)
                        rotate_uv(v[0], v[NF - 1], rot[k].x, rot[k].y);
                    if ((vmask >> k) & 1u) {
                        __stcs(base0 + off0 + poff(k), conv(v[0]));
                        if (NF == 2)
                            __stcs(base1 + off0 + poff(k), conv(v[NF - 1]));
                    }
                }
                base0 += g.out_level;
                if (NF == 2)
                    base1 += g.out_level;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------ gather, bulk-store form
// Same staging and arithmetic as k_gather_bilinear_staged, but the results of a batch of LZ levels are parked in a shared-memory
// OUTPUT tile [LZ][Tile::Y][Tile::X] and leave the SM through the bulk-copy engine (TMA), not through per-thread STG:
//   TENSOR = false: one cp.async.bulk.global.shared::cta (UBLKCP) per tile row and level, issued by LZ * Tile::Y threads;
//   TENSOR = true : one cp.async.bulk.tensor.3d (UTMASTG) per batch -- box Tile::X x Tile::Y x LZ over the [z][y][x] output,
//                   the tensor map clips rows / columns / levels outside the grid.
// Why (profiles/ubench/tma_store_bw.cu, 2000-wide rows, pure stores): a 64 x 16 tile written with per-thread 128-byte warp
// stores reaches 5.5 TB/s, the same tile through bulk copies 6.2-6.4 TB/s (the copy engine writes each 256-byte row run as one
// request stream, independent of the lane mapping); and the store leaves the instruction stream: 16 STS with immediate offsets
// per thread and step instead of 16 STG + their 64-bit address arithmetic (18 % of the instructions of the STG kernel).
// Tiles with more taps than the tap-major staging holds (pole, seam) are skipped here; the STG kernel runs on exactly those
// (TileTable::d_slow).
__device__ __forceinline__ void bulk_store_row(void* gmem, const void* smem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem), "r"((unsigned)__cvta_generic_to_shared(smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_box(const CUtensorMap* map, const void* smem, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
                 "r"((unsigned)__cvta_generic_to_shared(smem)), "r"(x), "r"(y), "r"(z)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all() // every bulk store this thread issued has finished READING shared memory
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_shared() // generic-proxy shared stores -> visible to the async proxy (the copy engine)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// four 32-bit values of one level into the shared-memory output tile with one 128-bit store
template <class T>
__device__ __forceinline__ void store_tile4(T* dst, T a, T b, T c, T d)
{
    static_assert(sizeof(T) == 4, "32-bit elements");
    uint4 w;
    w.x = *reinterpret_cast<unsigned*>(&a), w.y = *reinterpret_cast<unsigned*>(&b), w.z = *reinterpret_cast<unsigned*>(&c),
    w.w = *reinterpret_cast<unsigned*>(&d);
    *reinterpret_cast<uint4*>(dst) = w;
}

template <int LZ, bool TENSOR, class Out>
__global__ void __launch_bounds__(kThreads, LZ == 4 ? 3 : 2)
    k_gather_bilinear_bulk(const __grid_constant__ CUtensorMap omap, GatherGeom g, int tiles_x, const int* __restrict__ taps,
                           const int* __restrict__ ntaps_tab, const uint4* __restrict__ meta, const float4* __restrict__ xf4,
                           const float4* __restrict__ yf4, const float* __restrict__ in0, typename Out::type* __restrict__ out0, Out conv,
                           int fill_in, float bad0, int per)
{
    typedef typename Out::type T;
    typedef Tile<false> TL;
    constexpr int kStride = LZ == 8 ? kLvlStride : 4; // words between consecutive taps of the tap-major staging buffer
    static_assert(LZ == 4 || LZ == 8, "levels per batch");
    static_assert(LZ * TL::Y <= kThreads, "one row copy per thread");
    extern __shared__ __align__(128) float s_dyn[];
    float* const s_stage = s_dyn;                                             // [2][kStageFloats]
    T* const s_out = reinterpret_cast<T*>(s_dyn + 2 * kStageFloats);          // [2][LZ][TL::Y][TL::X]
    const int tile = blockIdx.x;
    const int t = threadIdx.x;
    const int ntaps = __ldg(ntaps_tab + tile);
    if (ntaps > kFastTaps)
        return; // whole CTA: a many-tap tile, done by the STG kernel
    const int* my_taps = taps + (size_t)tile * kMaxTaps;
    const size_t slot = (size_t)tile * kThreads + t;
    const uint4 m = __ldg(meta + slot);
    const float4 fx = __ldg(xf4 + slot), fy = __ldg(yf4 + slot);
    const unsigned mm[4] = {m.x, m.y, m.z, m.w};
    const float xf[4] = {fx.x, fx.y, fx.z, fx.w}, yf[4] = {fy.x, fy.y, fy.z, fy.w};
    int ia[4], ib[4], mode[4];
    float wx0[4], wy0[4];
    bool all_full = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ia[k] = (int)(mm[k] & 0xfffu) * kStride;
        ib[k] = (int)((mm[k] >> 12) & 0xfffu) * kStride;
        mode[k] = (int)(mm[k] >> 24);
        wx0[k] = __fsub_rn(1.f, xf[k]);
        wy0[k] = __fsub_rn(1.f, yf[k]);
        all_full = all_full && (mode[k] == FB_BL_FULL);
    }
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const long long z0 = (long long)blockIdx.y * per;
    const long long z1 = z0 + per < g.nz ? z0 + per : g.nz;
    if (z0 >= z1)
        return;
    // position of this thread's point 0 inside a level of the output tile; point k sits TL::RowStep rows further down
    const int spos = t;                  // == (t / TL::X) * TL::X + t % TL::X
    constexpr int kRowJump = kThreads;   // == TL::RowStep * TL::X
    // the row this thread copies out (TENSOR = false): level cr_l of the batch, tile row cr_y
    const int cr_l = t / TL::Y, cr_y = t % TL::Y;
    const int gx0 = tx * TL::X, gy = ty * TL::Y + cr_y;
    const int cols = g.ox - gx0 < TL::X ? g.ox - gx0 : TL::X;
    const bool copier = !TENSOR && t < LZ * TL::Y && gy < g.oy;
    const int tap0 = (t < ntaps) ? __ldg(my_taps + t) : -1;

    auto issue = [&](int buf, long long z, int nb) {
        float* dst = s_stage + buf * kStageFloats;
        const float* lv = in0 + z * g.in_level;
        if (tap0 >= 0) {
            const float* src = lv + tap0;
            float* d = dst + t * kStride;
#pragma unroll
            for (int zi = 0; zi < LZ; ++zi)
                if (zi < nb)
                    cp_async_f32(d + zi, src + zi * g.in_level);
        }
        if (ntaps > kThreads) {
            for (int r = t + kThreads; r < ntaps; r += kThreads) {
                const float* src = lv + __ldg(my_taps + r);
                for (int zi = 0; zi < nb; ++zi)
                    cp_async_f32(dst + r * kStride + zi, src + zi * g.in_level);
            }
        }
        cp_async_commit();
    };
    auto patch = [&](int buf, int nb) { // mifi_bad2nanf on the elements this thread copied itself
        const float nanv = undef_f();
        float* dst = s_stage + buf * kStageFloats;
        for (int r = t; r < ntaps; r += kThreads)
            for (int zi = 0; zi < nb; ++zi)
                if (dst[r * kStride + zi] == bad0)
                    dst[r * kStride + zi] = nanv;
    };
    auto send = [&](int buf, long long z, int nb) { // the finished output tile of batch (buf, z) leaves through the copy engine
        const T* src = s_out + (size_t)buf * LZ * kTilePts;
        if (TENSOR) {
            if (t == 0) {
                bulk_store_box(&omap, src, gx0, ty * TL::Y, (int)z);
                bulk_commit();
            }
        } else if (copier && cr_l < nb) {
            bulk_store_row(out0 + (z + cr_l) * g.out_level + (long long)gy * g.ox + gx0, src + (cr_l * TL::Y + cr_y) * TL::X,
                           (unsigned)(cols * sizeof(T)));
            bulk_commit();
        }
    };

    issue(0, z0, (int)((z1 - z0) < LZ ? (z1 - z0) : LZ));
    int buf = 0;
    long long zprev = z0;
    int nbprev = 0;
    for (long long z = z0; z < z1; z += LZ, buf ^= 1) {
        const int nb = (int)((z1 - z) < LZ ? (z1 - z) : LZ);
        cp_async_wait_pending<0>();
        if (fill_in)
            patch(buf, nb);
        bulk_wait_read_all(); // this thread's copies out of s_out[buf] (issued two batches ago) are done reading it
        __syncthreads();      // batch z is staged; everyone has finished (and fenced) the output tile of the previous batch
        if (z + LZ < z1)
            issue(buf ^ 1, z + LZ, (int)((z1 - z - LZ) < LZ ? (z1 - z - LZ) : LZ));
        if (nbprev)
            send(buf ^ 1, zprev, nbprev);
        const float* lvl = s_stage + buf * kStageFloats;
        T* so = s_out + (size_t)buf * LZ * kTilePts + spos;
        if (all_full) {
#pragma unroll
            for (int q = 0; q < LZ / 4; ++q) {
                float r[4][4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 a0 = *reinterpret_cast<const float4*>(lvl + ia[k] + 4 * q);
                    const float4 a1 = *reinterpret_cast<const float4*>(lvl + ia[k] + kStride + 4 * q);
                    const float4 b0 = *reinterpret_cast<const float4*>(lvl + ib[k] + 4 * q);
                    const float4 b1 = *reinterpret_cast<const float4*>(lvl + ib[k] + kStride + 4 * q);
                    r[k][0] = bilinear_full(wx0[k], xf[k], wy0[k], yf[k], a0.x, a1.x, b0.x, b1.x);
                    r[k][1] = bilinear_full(wx0[k], xf[k], wy0[k], yf[k], a0.y, a1.y, b0.y, b1.y);
                    r[k][2] = bilinear_full(wx0[k], xf[k], wy0[k], yf[k], a0.z, a1.z, b0.z, b1.z);
                    r[k][3] = bilinear_full(wx0[k], xf[k], wy0[k], yf[k], a0.w, a1.w, b0.w, b1.w);
                }
#pragma unroll
                for (int l = 0; l < 4; ++l)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        so[(4 * q + l) * kTilePts + k * kRowJump] = conv(r[k][l]);
            }
        } else { // a point on an edge strip or outside the source grid: per-point mode (interpolation.c:904-953)
            for (int zi = 0; zi < nb; ++zi) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float v = undef_f();
                    const float* qa = lvl + ia[k] + zi;
                    const float* qb = lvl + ib[k] + zi;
                    switch (mode[k]) {
                    case FB_BL_FULL:
                        v = bilinear_full(wx0[k], xf[k], wy0[k], yf[k], qa[0], qa[kStride], qb[0], qb[kStride]);
                        break;
                    case FB_BL_XLIN:
                        v = __fadd_rn(__fmul_rn(wx0[k], qa[0]), __fmul_rn(xf[k], qa[kStride]));
                        break;
                    case FB_BL_YLIN:
                        v = __fadd_rn(__fmul_rn(wy0[k], qa[0]), __fmul_rn(yf[k], qb[0]));
                        break;
                    case FB_BL_NEAR:
                        v = qa[0];
                        break;
                    default:
                        break;
                    }
                    so[zi * kTilePts + k * kRowJump] = conv(v);
                }
            }
        }
        fence_async_shared();
        zprev = z;
        nbprev = nb;
    }
    __syncthreads();
    send(buf ^ 1, zprev, nbprev);
    bulk_wait_read_all(); // shared memory must outlive the copies that read it
}


// ------------------------------------------------------------------------------------------------ nearest neighbour, bulk-store form
// The copied values of a batch of 8 levels are parked in a shared-memory output tile [8][8][128] (a thread's 4 x-neighbours are
// one 128-bit shared store per level) and leave through the copy engine as 128 x 8 x 8 boxes (TENSOR) or 512-byte row copies.
// Unlike the bilinear gather, nearest neighbour loads ONE tap per output, so the shared-memory pipe has room for the extra
// store and the copy engine's read (3 wavefronts per 32 outputs instead of 2), and the better store pattern shows:
// profiles/r02_nn_bulk_ab.txt.
template <bool TENSOR, class Out>
__global__ void __launch_bounds__(kThreads, 2)
    k_gather_nn_bulk(const __grid_constant__ CUtensorMap omap, GatherGeom g, int tiles_x, const int* __restrict__ taps,
                     const int* __restrict__ ntaps_tab, const uint4* __restrict__ meta, const float* __restrict__ in0,
                     typename Out::type* __restrict__ out0, Out conv, int fill_in, float bad0, int per)
{
    typedef typename Out::type T;
    typedef Tile<true> TL;
    constexpr int LZ = kMaxBatch;
    static_assert(sizeof(T) == 4, "32-bit output elements");
    static_assert(LZ * TL::Y <= kThreads, "one row copy per thread");
    extern __shared__ __align__(128) float s_dyn[];
    float* const s_stage = s_dyn;                                    // [2][kStageFloats]
    T* const s_out = reinterpret_cast<T*>(s_dyn + 2 * kStageFloats); // [2][LZ][TL::Y][TL::X]
    const int tile = blockIdx.x;
    const int t = threadIdx.x;
    const int ntaps = __ldg(ntaps_tab + tile);
    if (ntaps > kFastTaps)
        return; // a many-tap tile: done by the STG kernel (TileTable::d_slow)
    const int* my_taps = taps + (size_t)tile * kMaxTaps;
    const uint4 m = __ldg(meta + (size_t)tile * kThreads + t);
    const unsigned mm[4] = {m.x, m.y, m.z, m.w};
    int ia[4];
    bool hit[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ia[k] = (int)(mm[k] & 0xfffu) * kLvlStride;
        hit[k] = (int)(mm[k] >> 24) == FB_BL_NEAR; // else outside the source grid: MIFI_UNDEFINED_F (interpolation.c:872-875)
    }
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const long long z0 = (long long)blockIdx.y * per;
    const long long z1 = z0 + per < g.nz ? z0 + per : g.nz;
    if (z0 >= z1)
        return;
    const int cr_l = t / TL::Y, cr_y = t % TL::Y; // the row this thread copies out (!TENSOR)
    const int gx0 = tx * TL::X, gy = ty * TL::Y + cr_y;
    const int cols = g.ox - gx0 < TL::X ? g.ox - gx0 : TL::X;
    const bool copier = !TENSOR && t < LZ * TL::Y && gy < g.oy;
    const int tap0 = (t < ntaps) ? __ldg(my_taps + t) : -1;

    auto issue = [&](int buf, long long z, int nb) {
        float* dst = s_stage + buf * kStageFloats;
        const float* lv = in0 + z * g.in_level;
        if (tap0 >= 0) {
            const float* src = lv + tap0;
            float* d = dst + t * kLvlStride;
#pragma unroll
            for (int zi = 0; zi < LZ; ++zi)
                if (zi < nb)
                    cp_async_f32(d + zi, src + zi * g.in_level);
        }
        if (ntaps > kThreads) {
            for (int r = t + kThreads; r < ntaps; r += kThreads) {
                const float* src = lv + __ldg(my_taps + r);
                for (int zi = 0; zi < nb; ++zi)
                    cp_async_f32(dst + r * kLvlStride + zi, src + zi * g.in_level);
            }
        }
        cp_async_commit();
    };
    auto patch = [&](int buf, int nb) { // mifi_bad2nanf on the elements this thread copied itself
        const float nanv = undef_f();
        float* dst = s_stage + buf * kStageFloats;
        for (int r = t; r < ntaps; r += kThreads)
            for (int zi = 0; zi < nb; ++zi)
                if (dst[r * kLvlStride + zi] == bad0)
                    dst[r * kLvlStride + zi] = nanv;
    };
    auto send = [&](int buf, long long z, int nb) {
        const T* src = s_out + (size_t)buf * LZ * kTilePts;
        if (TENSOR) {
            if (t == 0) {
                bulk_store_box(&omap, src, gx0, ty * TL::Y, (int)z);
                bulk_commit();
            }
        } else if (copier && cr_l < nb) {
            bulk_store_row(out0 + (z + cr_l) * g.out_level + (long long)gy * g.ox + gx0, src + (cr_l * TL::Y + cr_y) * TL::X,
                           (unsigned)(cols * sizeof(T)));
            bulk_commit();
        }
    };

    issue(0, z0, (int)((z1 - z0) < LZ ? (z1 - z0) : LZ));
    int buf = 0;
    long long zprev = z0;
    int nbprev = 0;
    for (long long z = z0; z < z1; z += LZ, buf ^= 1) {
        const int nb = (int)((z1 - z) < LZ ? (z1 - z) : LZ);
        cp_async_wait_pending<0>();
        if (fill_in)
            patch(buf, nb);
        bulk_wait_read_all(); // this thread's copies out of s_out[buf] (issued two batches ago) are done reading it
        __syncthreads();      // batch z is staged; everyone has finished (and fenced) the output tile of the previous batch
        if (z + LZ < z1)
            issue(buf ^ 1, z + LZ, (int)((z1 - z - LZ) < LZ ? (z1 - z - LZ) : LZ));
        if (nbprev)
            send(buf ^ 1, zprev, nbprev);
        const float* lvl = s_stage + buf * kStageFloats;
        T* so = s_out + (size_t)buf * LZ * kTilePts + 4 * t; // this thread's 4 x-neighbours: tile positions 4t .. 4t+3 of every level
        const float nanv = undef_f();
#pragma unroll
        for (int q = 0; q < LZ / 4; ++q) {
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[k] = *reinterpret_cast<const float4*>(lvl + ia[k] + 4 * q); // four levels of the point's one tap
                if (!hit[k])
                    v[k] = make_float4(nanv, nanv, nanv, nanv);
            }
            // levels past nb hold stale values: written to the tile, never sent
            store_tile4(so + (4 * q + 0) * kTilePts, conv(v[0].x), conv(v[1].x), conv(v[2].x), conv(v[3].x));
            store_tile4(so + (4 * q + 1) * kTilePts, conv(v[0].y), conv(v[1].y), conv(v[2].y), conv(v[3].y));
            store_tile4(so + (4 * q + 2) * kTilePts, conv(v[0].z), conv(v[1].z), conv(v[2].z), conv(v[3].z));
            store_tile4(so + (4 * q + 3) * kTilePts, conv(v[0].w), conv(v[1].w), conv(v[2].w), conv(v[3].w));
        }
        fence_async_shared();
        zprev = z;
        nbprev = nb;
    }
    __syncthreads();
    send(buf ^ 1, zprev, nbprev);
    bulk_wait_read_all(); // shared memory must outlive the copies that read it
}

} // namespace

bool tile_table_supported(int ix, int iy, int ox, int oy)
{
    const long long in_level = (long long)ix * iy;
    const long long tiles = (long long)((ox + 63) / 64) * ((oy + 7) / 8); // at least as many as either tile shape needs
    return in_level > 0 && in_level < (1ll << 30) && ox > 0 && oy > 0 && tiles < 2147483647LL;
}

int tile_table_build(bool nn, const double* d_px, const double* d_py, int ix, int iy, int ox, int oy, TileTable* tt, cudaStream_t st)
{
    tile_table_free(tt);
    // bilinear: the column layout (lane = x, a thread owns 4 rows) unless FIMEX_B200_BILINEAR_QUAD=1 asks for the quad layout
    // (a thread owns 4 x-neighbours, taps reused in registers, 128-bit stores).  Opt-in after a same-box A/B on config 2
    // (profiles/r02_bilinear_quad_ab.txt): bit-identical, 22 % fewer shared-load wavefronts and the better store pattern, but
    // 12.6 ms against 10.9 ms -- the per-lane predicated re-loads and register moves raise the instruction count from 17.5 to
    // 25 per output and the kernel becomes issue-bound.
    const char* env = std::getenv("FIMEX_B200_BILINEAR_QUAD");
    const bool quad = nn || (env && env[0] == '1');
    const int tile_x = quad ? Tile<true>::X : Tile<false>::X, tile_y = quad ? Tile<true>::Y : Tile<false>::Y;
    tt->tiles_x = (ox + tile_x - 1) / tile_x;
    tt->tiles_y = (oy + tile_y - 1) / tile_y;
    const size_t tiles = (size_t)tt->tiles_x * tt->tiles_y;
    FB_CUDA_CHECK(cudaMalloc(&tt->d_cells, sizeof(int) * tiles * kMaxTaps));
    FB_CUDA_CHECK(cudaMalloc(&tt->d_ncells, sizeof(int) * tiles));
    FB_CUDA_CHECK(cudaMalloc(&tt->d_meta, sizeof(uint4) * tiles * kThreads));
    tt->nn = nn;
    tt->quad = quad;
    if (nn) {
        k_compile_tiles<true, true><<<(unsigned)tiles, kThreads, 0, st>>>(d_px, d_py, ox, oy, ix, iy, tt->tiles_x, tt->d_cells, tt->d_ncells,
                                                                          tt->d_meta, nullptr, nullptr);
    } else {
        FB_CUDA_CHECK(cudaMalloc(&tt->d_xf, sizeof(float4) * tiles * kThreads));
        FB_CUDA_CHECK(cudaMalloc(&tt->d_yf, sizeof(float4) * tiles * kThreads));
        if (quad)
            k_compile_tiles<false, true><<<(unsigned)tiles, kThreads, 0, st>>>(d_px, d_py, ox, oy, ix, iy, tt->tiles_x, tt->d_cells,
                                                                               tt->d_ncells, tt->d_meta, tt->d_xf, tt->d_yf);
        else
            k_compile_tiles<false, false><<<(unsigned)tiles, kThreads, 0, st>>>(d_px, d_py, ox, oy, ix, iy, tt->tiles_x, tt->d_cells,
                                                                                tt->d_ncells, tt->d_meta, tt->d_xf, tt->d_yf);
    }
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    // the tiles whose tap list does not fit the tap-major staging buffer (over the pole, across the seam): the bulk-store
    // gather leaves them to the STG kernel, which is launched on exactly this list
    std::vector<int> ntaps(tiles), slow;
    FB_CUDA_CHECK(cudaMemcpyAsync(ntaps.data(), tt->d_ncells, sizeof(int) * tiles, cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    for (size_t i = 0; i < tiles; ++i)
        if (ntaps[i] > kFastTaps)
            slow.push_back((int)i);
    tt->n_slow = (int)slow.size();
    if (!slow.empty()) {
        FB_CUDA_CHECK(cudaMalloc(&tt->d_slow, sizeof(int) * slow.size()));
        FB_CUDA_CHECK(cudaMemcpyAsync(tt->d_slow, slow.data(), sizeof(int) * slow.size(), cudaMemcpyHostToDevice, st));
        FB_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    return FB_OK;
}

void tile_table_free(TileTable* tt)
{
    if (tt->d_slow)
        cudaFree(tt->d_slow);
    if (tt->d_cells)
        cudaFree(tt->d_cells);
    if (tt->d_ncells)
        cudaFree(tt->d_ncells);
    if (tt->d_meta)
        cudaFree(tt->d_meta);
    if (tt->d_xf)
        cudaFree(tt->d_xf);
    if (tt->d_yf)
        cudaFree(tt->d_yf);
    *tt = TileTable();
}

namespace {
constexpr size_t kStageBytes = kBuffers * kStageFloats * sizeof(float); // per field

template <bool NN, class Out>
void launch_staged_as(dim3 grid, const GatherGeom& g, const TileTable& tt, const float* d_in, void* d_out, Out conv, const SliceConv& sc,
                      cudaStream_t st, const int* tile_list = nullptr)
{
    typedef typename Out::type T;
    const int vec_ok = ((g.ox % 4) == 0 && (reinterpret_cast<uintptr_t>(d_out) & (4 * sizeof(T) - 1)) == 0) ? 1 : 0;
    T* out = static_cast<T*>(d_out);
    const int fill_in = sc.fill_in ? 1 : 0;
    if (NN || tt.quad)
        k_gather_bilinear_staged<NN, 1, false, Out, true><<<grid, kThreads, kStageBytes, st>>>(g, tt.tiles_x, tt.d_cells, tt.d_ncells, tt.d_meta,
                                                                                                tt.d_xf, tt.d_yf, d_in, nullptr, out, nullptr, nullptr,
                                                                                                conv, fill_in, sc.bad_in[0], 0.f, vec_ok, tile_list);
    else
        k_gather_bilinear_staged<false, 1, false, Out, false><<<grid, kThreads, kStageBytes, st>>>(g, tt.tiles_x, tt.d_cells, tt.d_ncells, tt.d_meta,
                                                                                                    tt.d_xf, tt.d_yf, d_in, nullptr, out, nullptr,
                                                                                                    nullptr, conv, fill_in, sc.bad_in[0], 0.f, vec_ok,
                                                                                                    tile_list);
}

template <bool NN, bool ROT, bool QUAD>
cudaError_t launch_staged_vector_as(dim3 grid, const GatherGeom& g, const TileTable& tt, const double2* d_cs, const float* d_u, const float* d_v,
                                    float* d_uo, float* d_vo, const SliceConv& sc, cudaStream_t st)
{
    auto kernel = k_gather_bilinear_staged<NN, 2, ROT, StorePlain, QUAD>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * kStageBytes));
    if (e != cudaSuccess)
        return e;
    const uintptr_t align = reinterpret_cast<uintptr_t>(d_uo) | reinterpret_cast<uintptr_t>(d_vo);
    const int vec_ok = ((g.ox % 4) == 0 && (align & 15u) == 0) ? 1 : 0;
    kernel<<<grid, kThreads, 2 * kStageBytes, st>>>(g, tt.tiles_x, tt.d_cells, tt.d_ncells, tt.d_meta, tt.d_xf, tt.d_yf, d_u, d_v, d_uo, d_vo, d_cs,
                                                    StorePlain(), sc.fill_in ? 1 : 0, sc.bad_in[0], sc.bad_in[1], vec_ok, nullptr);
    return cudaSuccess;
}

template <bool NN, bool ROT>
cudaError_t launch_staged_vector(dim3 grid, const GatherGeom& g, const TileTable& tt, const double2* d_cs, const float* d_u, const float* d_v,
                                 float* d_uo, float* d_vo, const SliceConv& sc, cudaStream_t st)
{
    if (NN || tt.quad)
        return launch_staged_vector_as<NN, ROT, true>(grid, g, tt, d_cs, d_u, d_v, d_uo, d_vo, sc, st);
    return launch_staged_vector_as<false, ROT, false>(grid, g, tt, d_cs, d_u, d_v, d_uo, d_vo, sc, st);
}

template <bool NN>
bool launch_staged_typed(dim3 grid, const GatherGeom& g, const TileTable& tt, const float* d_in, void* d_out, const SliceConv& sc,
                         cudaStream_t st, const int* tile_list = nullptr)
{
    if (!sc.convert_out) {
        launch_staged_as<NN>(grid, g, tt, d_in, d_out, StorePlain(), sc, st, tile_list);
        return true;
    }
    switch (sc.out_type) {
#define FB_CASE(TAG, T)                                                                                                                    \
    case TAG:                                                                                                                              \
        launch_staged_as<NN>(grid, g, tt, d_in, d_out, StoreAs<T>{cast_fill<T>(sc.fill_out)}, sc, st, tile_list);                          \
        return true;
        FB_CASE(FB_T_FLOAT, float)
        FB_CASE(FB_T_DOUBLE, double)
        FB_CASE(FB_T_CHAR, signed char)
        FB_CASE(FB_T_SHORT, short)
        FB_CASE(FB_T_INT, int)
        FB_CASE(FB_T_UCHAR, unsigned char)
        FB_CASE(FB_T_USHORT, unsigned short)
        FB_CASE(FB_T_UINT, unsigned int)
#undef FB_CASE
    default:
        return false;
    }
}
} // namespace

// output types the staged kernels convert to while storing (the others go through a float slab and a cast pass)
bool staged_store_supports(int out_type)
{
    switch (out_type) {
    case FB_T_FLOAT:
    case FB_T_DOUBLE:
    case FB_T_CHAR:
    case FB_T_SHORT:
    case FB_T_INT:
    case FB_T_UCHAR:
    case FB_T_USHORT:
    case FB_T_UINT:
        return true;
    default:
        return false;
    }
}

namespace {

// cuTensorMapEncodeTiled through the runtime's driver entry point: the library does not link libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder()
{
    static const EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// 3-D map of the [z][y][x] output for boxes of Tile::X x Tile::Y x lz elements
bool output_tensor_map(void* d_out, const GatherGeom& g, size_t elem, int lz, CUtensorMap* map, bool quad = false)
{
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc)
        return false;
    const cuuint64_t dims[3] = {(cuuint64_t)g.ox, (cuuint64_t)g.oy, (cuuint64_t)g.nz};
    const cuuint64_t strides[2] = {(cuuint64_t)g.ox * elem, (cuuint64_t)g.out_level * elem};
    const cuuint32_t box[3] = {(cuuint32_t)(quad ? Tile<true>::X : Tile<false>::X), (cuuint32_t)(quad ? Tile<true>::Y : Tile<false>::Y),
                               (cuuint32_t)lz};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUtensorMapDataType dt = elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16; // moved as bits
    return enc(map, dt, 3, d_out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// FIMEX_B200_BULK_STORE: 0 = per-thread STG (the round-1 kernel), 1 = row copies (UBLKCP), 2 = tensor boxes (UTMASTG);
// FIMEX_B200_BULK_LEVELS: 4 or 8 levels per batch.  Read at every launch (A/B runs switch it inside one process).
int env_int(const char* name, int dflt)
{
    const char* e = std::getenv(name);
    return (e && *e) ? std::atoi(e) : dflt;
}

template <int LZ, bool TENSOR, class Out>
cudaError_t launch_bulk_as(dim3 grid, const CUtensorMap& map, const GatherGeom& g, const TileTable& tt, const float* d_in, void* d_out, Out conv,
                           const SliceConv& sc, int per, cudaStream_t st)
{
    typedef typename Out::type T;
    auto kernel = k_gather_bilinear_bulk<LZ, TENSOR, Out>;
    const size_t smem = 2 * kStageFloats * sizeof(float) + 2 * (size_t)LZ * kTilePts * sizeof(T);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    kernel<<<grid, kThreads, smem, st>>>(map, g, tt.tiles_x, tt.d_cells, tt.d_ncells, tt.d_meta, tt.d_xf, tt.d_yf, d_in, static_cast<T*>(d_out),
                                         conv, sc.fill_in ? 1 : 0, sc.bad_in[0], per);
    return cudaGetLastError();
}

template <class Out>
cudaError_t launch_bulk_shape(int lz, bool tensor, dim3 grid, const CUtensorMap& map, const GatherGeom& g, const TileTable& tt, const float* d_in,
                              void* d_out, Out conv, const SliceConv& sc, int per, cudaStream_t st)
{
    if (lz == 4)
        return tensor ? launch_bulk_as<4, true>(grid, map, g, tt, d_in, d_out, conv, sc, per, st)
                      : launch_bulk_as<4, false>(grid, map, g, tt, d_in, d_out, conv, sc, per, st);
    return tensor ? launch_bulk_as<8, true>(grid, map, g, tt, d_in, d_out, conv, sc, per, st)
                  : launch_bulk_as<8, false>(grid, map, g, tt, d_in, d_out, conv, sc, per, st);
}

} // namespace

int launch_gather_bilinear_staged(const GatherGeom& g, const TileTable& tt, const float* d_in, void* d_out, const SliceConv& sc,
                                  cudaStream_t st)
{
    if (g.out_level == 0 || g.nz == 0)
        return FB_OK;
    const unsigned tiles = (unsigned)tt.tiles_x * (unsigned)tt.tiles_y;
    // ---- nearest neighbour, 32-bit output elements: output tile in shared memory, stored by the copy engine ----
    // FIMEX_B200_NN_BULK: 0 = per-thread 128-bit stores, 1 = 512-byte row copies, 2 = tensor boxes 128 x 8 x 8
    {
        const int nn_mode = env_int("FIMEX_B200_NN_BULK", kNnBulkDefault);
        const bool t32 = !sc.convert_out || sc.out_type == FB_T_FLOAT || sc.out_type == FB_T_INT || sc.out_type == FB_T_UINT;
        if (tt.nn && nn_mode != 0 && t32 && (g.ox % 4) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0 && g.nz < 2147483647LL) {
            const bool tensor = nn_mode == 2;
            const int chunks = z_chunks(tiles, g.nz, env_int("FIMEX_B200_BULK_CHUNK", 64));
            int per = (int)((g.nz + chunks - 1) / chunks);
            per = (per + kMaxBatch - 1) / kMaxBatch * kMaxBatch; // whole batches per chunk: a tensor box never reaches into the next chunk
            dim3 grid(tiles, (unsigned)((g.nz + per - 1) / per));
            CUtensorMap map;
            std::memset(&map, 0, sizeof(map));
            FB_REQUIRE(!tensor || output_tensor_map(d_out, g, 4, kMaxBatch, &map, true), "cuTensorMapEncodeTiled failed for the output tensor");
            cudaError_t e = cudaErrorInvalidValue;
            const size_t smem = 2 * kStageFloats * sizeof(float) + 2 * (size_t)kMaxBatch * kTilePts * 4;
            auto go = [&](auto conv) -> cudaError_t {
                typedef decltype(conv) C;
                typedef typename C::type T;
                cudaError_t err;
                if (tensor) {
                    auto kernel = k_gather_nn_bulk<true, C>;
                    if ((err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
                        return err;
                    kernel<<<grid, kThreads, smem, st>>>(map, g, tt.tiles_x, tt.d_cells, tt.d_ncells, tt.d_meta, d_in, static_cast<T*>(d_out), conv,
                                                         sc.fill_in ? 1 : 0, sc.bad_in[0], per);
                } else {
                    auto kernel = k_gather_nn_bulk<false, C>;
                    if ((err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
                        return err;
                    kernel<<<grid, kThreads, smem, st>>>(map, g, tt.tiles_x, tt.d_cells, tt.d_ncells, tt.d_meta, d_in, static_cast<T*>(d_out), conv,
                                                         sc.fill_in ? 1 : 0, sc.bad_in[0], per);
                }
                return cudaGetLastError();
            };
            if (!sc.convert_out)
                e = go(StorePlain());
            else if (sc.out_type == FB_T_FLOAT)
                e = go(StoreAs<float>{cast_fill<float>(sc.fill_out)});
            else if (sc.out_type == FB_T_INT)
                e = go(StoreAs<int>{cast_fill<int>(sc.fill_out)});
            else
                e = go(StoreAs<unsigned int>{cast_fill<unsigned int>(sc.fill_out)});
            FB_CUDA_CHECK(e);
            count_launch();
            if (tt.n_slow > 0) { // the many-tap tiles, through the STG kernel
                dim3 sgrid((unsigned)tt.n_slow, z_chunks(tt.n_slow, g.nz, 64));
                FB_REQUIRE(launch_staged_typed<true>(sgrid, g, tt, d_in, d_out, sc, st, tt.d_slow), "staged gather: unsupported output type");
                count_launch();
                FB_CUDA_CHECK(cudaGetLastError());
            }
            return FB_OK;
        }
    }
    // ---- bilinear, 2- and 4-byte output elements: output tiles staged in shared memory, stored by the copy engine ----
    // Opt-in (default 0): measured on config 2, same box (profiles/r02_bulk_store_ab.txt), the bulk-store forms are bit-identical
    // but SLOWER than per-thread STG -- 12.8 ms (tensor boxes, 8 levels) / 14.6 ms (row copies) against 10.9 ms -- although the
    // same tiles store faster in isolation (6.2-6.4 against 5.5 TB/s): the gather is bound by the shared-memory data pipe, and the
    // output tile adds one STS and one copy-engine read of every value to the 4.45 tap-load wavefronts per 32 outputs.
    const int mode = env_int("FIMEX_B200_BULK_STORE", 0);
    size_t elem = 4;
    bool bulk_type = !sc.convert_out || sc.out_type == FB_T_FLOAT;
    if (sc.convert_out && (sc.out_type == FB_T_SHORT || sc.out_type == FB_T_USHORT)) {
        elem = 2;
        bulk_type = true;
    }
    if (sc.convert_out && (sc.out_type == FB_T_INT || sc.out_type == FB_T_UINT))
        bulk_type = true;
    const bool bulk_ok = !tt.nn && !tt.quad && mode != 0 && bulk_type && (g.ox * elem) % 16 == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0 &&
                         (g.out_level * (long long)elem) % 16 == 0 && g.nz < 2147483647LL;
    if (bulk_ok) {
        const int lz = env_int("FIMEX_B200_BULK_LEVELS", 8) == 4 ? 4 : 8;
        const bool tensor = mode == 2;
        const int chunks = z_chunks(tiles, g.nz, env_int("FIMEX_B200_BULK_CHUNK", 128));
        int per = (int)((g.nz + chunks - 1) / chunks);
        per = (per + lz - 1) / lz * lz; // whole batches per chunk: a tensor box never reaches into the next chunk's levels
        dim3 grid(tiles, (unsigned)((g.nz + per - 1) / per));
        CUtensorMap map;
        std::memset(&map, 0, sizeof(map));
        FB_REQUIRE(!tensor || output_tensor_map(d_out, g, elem, lz, &map), "cuTensorMapEncodeTiled failed for the output tensor");
        cudaError_t e = cudaErrorInvalidValue;
        if (!sc.convert_out) {
            e = launch_bulk_shape(lz, tensor, grid, map, g, tt, d_in, d_out, StorePlain(), sc, per, st);
        } else {
            switch (sc.out_type) {
#define FB_CASE(TAG, T)                                                                                                                    \
    case TAG:                                                                                                                              \
        e = launch_bulk_shape(lz, tensor, grid, map, g, tt, d_in, d_out, StoreAs<T>{cast_fill<T>(sc.fill_out)}, sc, per, st);               \
        break;
                FB_CASE(FB_T_FLOAT, float)
                FB_CASE(FB_T_SHORT, short)
                FB_CASE(FB_T_USHORT, unsigned short)
                FB_CASE(FB_T_INT, int)
                FB_CASE(FB_T_UINT, unsigned int)
#undef FB_CASE
            default:
                break;
            }
        }
        FB_CUDA_CHECK(e);
        count_launch();
        if (tt.n_slow > 0) { // the many-tap tiles, through the STG kernel
            dim3 sgrid((unsigned)tt.n_slow, z_chunks(tt.n_slow, g.nz, 128));
            FB_REQUIRE(launch_staged_typed<false>(sgrid, g, tt, d_in, d_out, sc, st, tt.d_slow), "staged gather: unsupported output type");
            count_launch();
            FB_CUDA_CHECK(cudaGetLastError());
        }
        return FB_OK;
    }
    // levels per CTA: same-box sweep on config 2 (ms per step), bilinear 64 / 96 / 128 / 160 -> 11.47 / 11.37 / 11.27 / 11.33,
    // nearest neighbour 48 / 64 / 96 -> 9.87 / 9.73 / 9.80
    dim3 grid(tiles, z_chunks(tiles, g.nz, tt.nn ? 64 : 128));
    const bool ok = tt.nn ? launch_staged_typed<true>(grid, g, tt, d_in, d_out, sc, st) : launch_staged_typed<false>(grid, g, tt, d_in, d_out, sc, st);
    FB_REQUIRE(ok, "staged gather: unsupported output type");
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

// both components of a vector through the staged gather (plain float output), rotated when d_cs != nullptr
int launch_gather_staged_vector(const GatherGeom& g, const TileTable& tt, const double2* d_cs, const float* d_u, const float* d_v, float* d_uo,
                                float* d_vo, const SliceConv& sc, cudaStream_t st)
{
    if (g.out_level == 0 || g.nz == 0)
        return FB_OK;
    FB_REQUIRE(!sc.convert_out, "staged vector gather writes plain floats");
    const unsigned tiles = (unsigned)tt.tiles_x * (unsigned)tt.tiles_y;
    dim3 grid(tiles, z_chunks(tiles, g.nz));
    cudaError_t e;
    if (tt.nn)
        e = d_cs ? launch_staged_vector<true, true>(grid, g, tt, d_cs, d_u, d_v, d_uo, d_vo, sc, st)
                 : launch_staged_vector<true, false>(grid, g, tt, d_cs, d_u, d_v, d_uo, d_vo, sc, st);
    else
        e = d_cs ? launch_staged_vector<false, true>(grid, g, tt, d_cs, d_u, d_v, d_uo, d_vo, sc, st)
                 : launch_staged_vector<false, false>(grid, g, tt, d_cs, d_u, d_v, d_uo, d_vo, sc, st);
    FB_CUDA_CHECK(e);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

} // namespace fb
