// fimex_b200/csrc/staged_kernels.cu -- K4 (bilinear) as a shared-memory staged gather: the fast path of
// CachedInterpolation::interpolateValues (/root/reference/src/CachedInterpolation.cc:118-147 with
// mifi_get_values_bilinear_f, src/interpolation.c:881-957).
//
// Why: the direct gather (gather_kernels.cu) issues four 4-byte global loads per output value and is bound by
// L1/LSU request rate and load latency (ncu, profiles/): ~33 % of the HBM roofline.  The target grid is
// normally finer than the source, so the 1024 target points of a tile touch only a few dozen distinct source
// cells.  Here a CTA owns a tile of 128 x 8 target points and, per batch of levels,
//   1. stages the tile's DISTINCT source cells into shared memory, one float4 {s00, s01, s10, s11} per cell,
//      with coalesced loads over a sorted cell list (each source value is requested once per tile, not once per
//      target point);
//   2. every thread then needs ONE 128-bit shared-memory load per output value, evaluates the reference's
//      formula (no FMA contraction: bit-identical to the CPU), and streams its 4 consecutive outputs with one
//      128-bit st.global.cs, a warp writing 512 contiguous bytes of a row.
// The cell list is per tile, not a bounding box, so tiles over the pole or across the 0/360 longitude seam
// (where the footprint of a tile is scattered) cost no more than their number of distinct cells.
//
// Tile table (built once per grid by k_compile_tiles, on the device):
//   cells [tile][1024]  int   sorted distinct cell offsets (y0*ix + x0), bit 30 = has a right neighbour,
//                             bit 31 = has a lower neighbour (taps outside the level are never loaded)
//   ncells[tile]        int
//   meta  [tile][256]   uint4 per thread: 4 x (local cell index | mode << 16)
//   xf/yf [tile][256]   float4 per thread: the reference's float xfrac / yfrac of its 4 points
#include "kernels.h"
#include "tables.cuh"

namespace fb {

namespace {

constexpr int kThreads = 256;
constexpr int kTileX = 128, kTileY = 8, kTilePts = kTileX * kTileY; // 1024 points, 4 consecutive x per thread
constexpr int kStageCells = 1024;                                    // float4 slots of staging memory (16 KB)
constexpr int kMaxBatch = 8;                                         // levels staged per barrier pair
constexpr unsigned kHasRight = 1u << 30, kHasDown = 1u << 31, kOffMask = (1u << 30) - 1;
constexpr int kNoCell = 0x7fffffff;

// ------------------------------------------------------------------------------------------------ table compiler
__global__ void __launch_bounds__(kThreads) k_compile_tiles(const double* __restrict__ px, const double* __restrict__ py, int ox, int oy,
                                                          int ix, int iy, int tiles_x, int* __restrict__ cells, int* __restrict__ ncells,
                                                          uint4* __restrict__ meta, float4* __restrict__ xf4, float4* __restrict__ yf4)
{
    __shared__ int s_keys[kTilePts];
    __shared__ int s_uniq[kTilePts];
    __shared__ int s_warp_tot[kThreads / 32];
    const int tile = blockIdx.x;
    const int t = threadIdx.x;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int y = ty * kTileY + (t >> 5);
    const int x0 = tx * kTileX + (t & 31) * 4;
    int off[4], mode[4];
    float xf[4], yf[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int4 e = make_int4(0, 0, 0, FB_BL_NAN);
        if (y < oy && x0 + k < ox) {
            const long long i = (long long)y * ox + x0 + k;
            e = classify_bilinear(px[i], py[i], ix, iy);
        }
        off[k] = e.x;
        xf[k] = __int_as_float(e.y);
        yf[k] = __int_as_float(e.z);
        mode[k] = e.w;
        s_keys[t * 4 + k] = (e.w == FB_BL_NAN) ? kNoCell : e.x;
    }
    __syncthreads();
    // bitonic sort of the 1024 keys, ascending
    for (int k = 2; k <= kTilePts; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < kTilePts; i += kThreads) {
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
    // distinct keys -> s_uniq (block-wide exclusive scan of the per-thread counts)
    int flags = 0, cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = t * 4 + k;
        const int v = s_keys[i];
        const bool first = (v != kNoCell) && (i == 0 || s_keys[i - 1] != v);
        flags |= first ? (1 << k) : 0;
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
    for (int k = 0; k < 4; ++k)
        if (flags & (1 << k))
            s_uniq[pos++] = s_keys[t * 4 + k];
    __syncthreads();
    // publish the cell list with its neighbour flags
    for (int j = t; j < total; j += kThreads) {
        const int o = s_uniq[j];
        const int cx = o % ix, cy = o / ix;
        unsigned packed = (unsigned)o;
        if (cx + 1 < ix)
            packed |= kHasRight;
        if (cy + 1 < iy)
            packed |= kHasDown;
        cells[(size_t)tile * kTilePts + j] = (int)packed;
    }
    if (t == 0)
        ncells[tile] = total;
    // each point looks its cell up in the sorted list
    unsigned m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int local = 0;
        if (mode[k] != FB_BL_NAN) {
            int lo = 0, hi = total - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_uniq[mid] < off[k])
                    lo = mid + 1;
                else
                    hi = mid;
            }
            local = lo;
        }
        m[k] = (unsigned)local | ((unsigned)mode[k] << 16);
    }
    const size_t slot = (size_t)tile * kThreads + t;
    meta[slot] = make_uint4(m[0], m[1], m[2], m[3]);
    xf4[slot] = make_float4(xf[0], xf[1], xf[2], xf[3]);
    yf4[slot] = make_float4(yf[0], yf[1], yf[2], yf[3]);
}

// ------------------------------------------------------------------------------------------------ gather
__device__ __forceinline__ float eval_cell(int mode, const float4 s, float wx0, float xf, float wy0, float yf)
{
    // s = {s00, s01, s10, s11}; formulas of interpolation.c:899-900, :911, :931, :940
    switch (mode) {
    case FB_BL_FULL: {
        const float top = __fadd_rn(__fmul_rn(wx0, s.x), __fmul_rn(xf, s.y));
        const float bot = __fadd_rn(__fmul_rn(wx0, s.z), __fmul_rn(xf, s.w));
        return __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(yf, bot));
    }
    case FB_BL_XLIN:
        return __fadd_rn(__fmul_rn(wx0, s.x), __fmul_rn(xf, s.y));
    case FB_BL_YLIN:
        return __fadd_rn(__fmul_rn(wy0, s.x), __fmul_rn(yf, s.z));
    case FB_BL_NEAR:
        return s.x;
    default:
        return undef_f();
    }
}

// 4-byte asynchronous global->shared copy (LDGSTS); src_bytes == 0 writes a zero without reading
__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gmem_src, int src_bytes)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

constexpr int kSlotNone = -2, kSlotZero = -1;

// offset (inside a level) of staging element r = 4*cell + tap, kSlotZero for a tap outside the level,
// kSlotNone past the end of the tile's list
__device__ __forceinline__ int staging_slot(const int* s_cells, int r, int ntaps, int ix)
{
    if (r >= ntaps)
        return kSlotNone;
    const unsigned packed = (unsigned)s_cells[r >> 2];
    const int tap = r & 3;
    const bool right = (packed & kHasRight) != 0, down = (packed & kHasDown) != 0;
    const bool ok = (!(tap & 1) || right) && (!(tap & 2) || down);
    return ok ? (int)(packed & kOffMask) + (tap & 1) + ((tap & 2) ? ix : 0) : kSlotZero;
}

template <bool VEC_STORE>
__global__ void __launch_bounds__(kThreads, 5) k_gather_bilinear_staged(GatherGeom g, int tiles_x, const int* __restrict__ cells,
                                                                     const int* __restrict__ ncells, const uint4* __restrict__ meta,
                                                                     const float4* __restrict__ xf4, const float4* __restrict__ yf4,
                                                                     const float* __restrict__ in, float* __restrict__ out)
{
    __shared__ float4 s_stage[2][kStageCells]; // double buffer: batch b+1 lands while batch b is consumed
    __shared__ int s_cells[kTilePts];
    const int tile = blockIdx.x;
    const int t = threadIdx.x;
    const int nc = __ldg(ncells + tile);
    for (int j = t; j < nc; j += kThreads)
        s_cells[j] = __ldg(cells + (size_t)tile * kTilePts + j);
    const size_t slot = (size_t)tile * kThreads + t;
    const uint4 m = __ldg(meta + slot);
    const float4 fx = __ldg(xf4 + slot), fy = __ldg(yf4 + slot);
    const unsigned mm[4] = {m.x, m.y, m.z, m.w};
    const float xf[4] = {fx.x, fx.y, fx.z, fx.w}, yf[4] = {fy.x, fy.y, fy.z, fy.w};
    int idx[4], mode[4];
    float wx0[4], wy0[4];
    bool all_full = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        idx[k] = (int)(mm[k] & 0xffffu);
        mode[k] = (int)(mm[k] >> 16);
        wx0[k] = __fsub_rn(1.f, xf[k]);
        wy0[k] = __fsub_rn(1.f, yf[k]);
        all_full = all_full && (mode[k] == FB_BL_FULL);
    }
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int y = ty * kTileY + (t >> 5);
    const int x0 = tx * kTileX + (t & 31) * 4;
    int valid = (y < g.oy) ? g.ox - x0 : 0;
    valid = valid > 4 ? 4 : (valid < 0 ? 0 : valid);
    const long long per = (g.nz + gridDim.y - 1) / gridDim.y;
    const long long z0 = (long long)blockIdx.y * per;
    const long long z1 = z0 + per < g.nz ? z0 + per : g.nz;
    float* o = out + z0 * g.out_level + (long long)y * g.ox + x0;
    const int zb = (nc <= kStageCells / kMaxBatch) ? kMaxBatch : (nc > 0 ? kStageCells / nc : kMaxBatch);
    const int ntaps = nc * 4;
    __syncthreads();
    // the staging elements this thread copies every level: r = t and r = t + 256 (covers tiles of <= 128 cells,
    // 96 % of them at the BASELINE geometry); larger tiles loop over the rest
    const int slot0 = staging_slot(s_cells, t, ntaps, g.ix);
    const int slot1 = staging_slot(s_cells, t + kThreads, ntaps, g.ix);

    auto issue = [&](int buf, long long z, int nb) {
        float* dst = reinterpret_cast<float*>(s_stage[buf]);
        const float* lv = in + z * g.in_level;
#pragma unroll 4
        for (int zi = 0; zi < nb; ++zi, lv += g.in_level, dst += ntaps) {
            if (slot0 != kSlotNone)
                cp_async_f32(dst + t, lv + (slot0 >= 0 ? slot0 : 0), slot0 >= 0 ? 4 : 0);
            if (slot1 != kSlotNone)
                cp_async_f32(dst + t + kThreads, lv + (slot1 >= 0 ? slot1 : 0), slot1 >= 0 ? 4 : 0);
            for (int r = t + 2 * kThreads; r < ntaps; r += kThreads) {
                const int s = staging_slot(s_cells, r, ntaps, g.ix);
                cp_async_f32(dst + r, lv + (s >= 0 ? s : 0), s >= 0 ? 4 : 0);
            }
        }
        cp_async_commit();
    };

    if (z0 < z1)
        issue(0, z0, (int)((z1 - z0) < zb ? (z1 - z0) : zb));
    int buf = 0;
    for (long long z = z0; z < z1; z += zb, buf ^= 1) {
        const int nb = (int)((z1 - z) < zb ? (z1 - z) : zb);
        cp_async_wait_all();
        __syncthreads(); // batch z has landed for every thread, and every thread is done reading the other buffer
        const long long zn = z + zb;
        if (zn < z1)
            issue(buf ^ 1, zn, (int)((z1 - zn) < zb ? (z1 - zn) : zb));
        if (valid == 0)
            continue;
        const float4* cell = s_stage[buf];
        if (all_full) {
#pragma unroll 2
            for (int zi = 0; zi < nb; ++zi, cell += nc) {
                float r[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 s = cell[idx[k]];
                    const float top = __fadd_rn(__fmul_rn(wx0[k], s.x), __fmul_rn(xf[k], s.y));
                    const float bot = __fadd_rn(__fmul_rn(wx0[k], s.z), __fmul_rn(xf[k], s.w));
                    r[k] = __fadd_rn(__fmul_rn(wy0[k], top), __fmul_rn(yf[k], bot));
                }
                if (VEC_STORE && valid == 4) {
                    __stcs(reinterpret_cast<float4*>(o), make_float4(r[0], r[1], r[2], r[3]));
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < valid)
                            __stcs(o + k, r[k]);
                }
                o += g.out_level;
            }
        } else {
            for (int zi = 0; zi < nb; ++zi, cell += nc) {
                float r[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 s = (mode[k] != FB_BL_NAN) ? cell[idx[k]] : make_float4(0.f, 0.f, 0.f, 0.f);
                    r[k] = eval_cell(mode[k], s, wx0[k], xf[k], wy0[k], yf[k]);
                }
                if (VEC_STORE && valid == 4) {
                    __stcs(reinterpret_cast<float4*>(o), make_float4(r[0], r[1], r[2], r[3]));
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < valid)
                            __stcs(o + k, r[k]);
                }
                o += g.out_level;
            }
        }
    }
}

int z_chunks_for(long long ctas_x, long long nz)
{
    // enough CTAs for ~16 waves of 8 per SM, but chunks of >= 32 levels so the per-chunk table read stays small
    const long long want = (long long)sm_count() * 8 * 16;
    long long gy = (want + ctas_x - 1) / ctas_x;
    const long long max_gy = nz / 32 > 1 ? nz / 32 : 1;
    if (gy > max_gy)
        gy = max_gy;
    if (gy < 1)
        gy = 1;
    if (gy > 65535)
        gy = 65535;
    return (int)gy;
}

} // namespace

bool tile_table_supported(int ix, int iy, int ox, int oy)
{
    const long long in_level = (long long)ix * iy;
    const long long tiles = (long long)((ox + kTileX - 1) / kTileX) * ((oy + kTileY - 1) / kTileY);
    return in_level > 0 && in_level < (1ll << 30) && ox > 0 && oy > 0 && tiles < 2147483647LL;
}

int tile_table_build(const double* d_px, const double* d_py, int ix, int iy, int ox, int oy, TileTable* tt, cudaStream_t st)
{
    tile_table_free(tt);
    tt->tiles_x = (ox + kTileX - 1) / kTileX;
    tt->tiles_y = (oy + kTileY - 1) / kTileY;
    const size_t tiles = (size_t)tt->tiles_x * tt->tiles_y;
    FB_CUDA_CHECK(cudaMalloc(&tt->d_cells, sizeof(int) * tiles * kTilePts));
    FB_CUDA_CHECK(cudaMalloc(&tt->d_ncells, sizeof(int) * tiles));
    FB_CUDA_CHECK(cudaMalloc(&tt->d_meta, sizeof(uint4) * tiles * kThreads));
    FB_CUDA_CHECK(cudaMalloc(&tt->d_xf, sizeof(float4) * tiles * kThreads));
    FB_CUDA_CHECK(cudaMalloc(&tt->d_yf, sizeof(float4) * tiles * kThreads));
    k_compile_tiles<<<(unsigned)tiles, kThreads, 0, st>>>(d_px, d_py, ox, oy, ix, iy, tt->tiles_x, tt->d_cells, tt->d_ncells, tt->d_meta,
                                                          tt->d_xf, tt->d_yf);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

void tile_table_free(TileTable* tt)
{
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

int launch_gather_bilinear_staged(const GatherGeom& g, const TileTable& tt, const float* d_in, float* d_out, cudaStream_t st)
{
    if (g.out_level == 0 || g.nz == 0)
        return FB_OK;
    const unsigned tiles = (unsigned)tt.tiles_x * (unsigned)tt.tiles_y;
    dim3 grid(tiles, z_chunks_for(tiles, g.nz));
    const bool vec = (g.ox % 4) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0;
    if (vec)
        k_gather_bilinear_staged<true><<<grid, kThreads, 0, st>>>(g, tt.tiles_x, tt.d_cells, tt.d_ncells, tt.d_meta, tt.d_xf, tt.d_yf, d_in,
                                                                  d_out);
    else
        k_gather_bilinear_staged<false><<<grid, kThreads, 0, st>>>(g, tt.tiles_x, tt.d_cells, tt.d_ncells, tt.d_meta, tt.d_xf, tt.d_yf, d_in,
                                                                   d_out);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

} // namespace fb
