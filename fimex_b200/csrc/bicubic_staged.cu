// fimex_b200/csrc/bicubic_staged.cu -- K5 (bicubic) as a shared-memory staged gather with register reuse of the
// 4x4 stencil: the fast path of CachedInterpolation::interpolateValues with mifi_get_values_bicubic_f
// (/root/reference/src/CachedInterpolation.cc:118-147, src/interpolation.c:959-1028), optionally for both
// components of a vector with the rotation of mifi_vector_reproject_values_by_matrix_f (:790-812) in the epilogue.
//
// What bounds bicubic on B200 (ncu, profiles/): not HBM.  The reference's arithmetic is 16 fp64 multiply-adds for the
// row sums plus 4 for the column sum per output, each multiply and add rounded separately (no FMA on x86-64), with the
// accumulator re-rounded to fp32 after every row.  Bit-exact replay costs 35 fp64 instructions and 7 fp32<->fp64
// conversions per output; the fp64 pipe issues 64 lanes/clk/SM and the conversion (XU) pipe 16 lanes/clk/SM.  The
// direct kernel (gather_kernels.cu) additionally converts its 16 taps per output on the XU pipe, which saturates it
// (92 % busy, 12 % of the HBM roofline).  So the design goal here is: no per-output tap conversion and as few
// shared-memory reads per output as possible, leaving the fp64 pipe as the only limiter (~0.55 clk per output and SM).
//
//   * a CTA owns a tile of 32 x 28 target points and a chunk of levels;
//   * the tile's DISTINCT source values ("taps") of a batch of levels are loaded once into registers, converted to
//     fp64 ONCE (per tap, not per use) and parked in shared memory, double buffered: the global loads of batch b+2 are
//     in flight while batch b+1 waits in shared memory and batch b is consumed;
//   * target points are regrouped by source cell: a thread owns a GROUP of up to 4 target points that read the SAME
//     4x4 stencil, loads each stencil row once (4 x LDS.64) and uses it for its 4 points -- "reuse of the stencil" is
//     done in registers rather than by warp shuffles, because the points of one cell are not adjacent lanes on a rotated
//     grid; per-point weights (fp64) stay in registers for the whole chunk;
//   * results go through a (double-buffered, bank-swizzled) shared-memory output tile so that global stores are full
//     128-byte rows (st.global.cs) and one barrier per batch suffices.
//
// Table (built once per grid on the device by k_compile_bicubic_tiles, two passes: count, then fill):
//   info  [tile]     int4   {first tap, ntaps (-1: direct fallback), first group, ngroups}
//   taps  [...]      int    per tile: sorted distinct source offsets y*ix + x
//   gmeta [group]    uint4  8 x u16: list index of the first tap of each of the 4 stencil rows, then the 4 point slots
//                           (out_slot(y*32 + x) inside the tile; 896 = padding)
//   gfrac [group][4] double2 (xfrac, yfrac) of the 4 points, fp64 as in the reference (:970-973)
// Tiles whose tap list does not fit the staging buffers (over the pole, where one tile sees thousands of source
// columns) are computed by the same kernel with direct global loads.
#include "kernels.h"
#include "interp_math.cuh"
#include "convert.cuh"

#include <cuda.h> // CUtensorMap (types only; the encoder comes through cudaGetDriverEntryPoint, libcuda is not linked)

#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

namespace fb {

namespace {

constexpr int kT = 256;                              // threads per CTA
constexpr int kTX = 32, kTY = 28, kTP = kTX * kTY;   // tile of 896 target points: 224 groups of 4 + padding <= 256
constexpr int kSortN = 1024;                         // tile points padded to a power of two for the bitonic sort
constexpr int kDump = kTP;                           // output-tile slot of padding points
constexpr int kOutRow = kTP + 4;                     // floats per (field, level) of the output tile (multiple of 4)
constexpr int kOutRowTma = 1024;                     // the same when the tile leaves through the copy engine: every (field, level) row on its own
                                                     // 4096-byte boundary, so that out_slot() IS the 128-byte swizzle of a tensor-map store
constexpr int kOutRows = 8;                          // (field, level) rows of the output tile
constexpr int kStageElems = 8 * kT;                  // tap values staged per batch: 8 registers per thread
constexpr int kTapCap = kStageElems;                 // tiles with more distinct taps are computed directly
constexpr int kFastTaps = kT;                        // at most this many taps: one tap per thread, 8 levels per batch
constexpr int kStageDoubles = 2 * kFastTaps * 5;     // per buffer: max(256*9, 2*256*5, 2048) = 2560 doubles
constexpr int kStageFloats32 = kFastTaps * 12;       // fp32 mode, per buffer: max(256*12, 2*256*4, 2048) = 3072 floats
constexpr int stage_elems(bool fp32) { return fp32 ? kStageFloats32 : kStageDoubles; }
constexpr int kMaxTapKeys = 16384;                   // 16 stencil taps for each of at most 896 distinct cells, padded to 2^k
constexpr int kNoKey = 0x7fffffff;
#ifndef FB_BIC_UNROLL
#define FB_BIC_UNROLL 1 // unroll factor of the level loop of a full batch (experiments: -DFB_BIC_UNROLL=4 makes every shared address an immediate)
#endif
constexpr int kBicUnroll = FB_BIC_UNROLL;
#ifndef FB_BIC_NL
#define FB_BIC_NL 2 // levels per inner step of a full batch (independent accumulator chains: 4 points x NL levels)
#endif
constexpr int kNL = FB_BIC_NL;

// output-tile slot of tile point p = y*32 + x: 4-float column groups are XOR-permuted by the row, so that the points of
// one source cell (a block of neighbouring rows and columns) spread over the banks; rows stay readable as float4.
// (What is left is the conflict rate of 32 effectively random banks: 2.8 wavefronts per store instruction in a simulation of
// config 2's tiles, 2.6 measured, and 2.8-3.0 for every other FIXED permutation tried -- xor or add of k * row over all five
// column bits, rotating the point order inside a group by the group index.  What does help is scheduling the point order
// per warp against the banks actually taken: step G of the table compiler)
__host__ __device__ __forceinline__ int out_slot(int p)
{
    return (p & ~31) | ((p & 31) ^ (((p >> 5) & 7) << 2));
}

// ------------------------------------------------------------------------------------------------ block helpers
// bitonic sort of n = 2^k keys in shared memory, ascending
template <class K>
__device__ __forceinline__ void bitonic_sort(K* s, int n)
{
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += kT) {
                const int p = i ^ j;
                if (p > i) {
                    const K a = s[i], b = s[p];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        s[i] = b;
                        s[p] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

struct OpAdd {
    __device__ __forceinline__ int operator()(int a, int b) const { return a + b; }
};
struct OpMax {
    __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; }
};

// inclusive scan over kT * N values (thread t owns elements N*t .. N*t+N-1); every thread gets the total
template <int N, class Op>
__device__ __forceinline__ int block_scan(int (&v)[N], Op op, int identity, int* s_warp)
{
#pragma unroll
    for (int k = 1; k < N; ++k)
        v[k] = op(v[k - 1], v[k]);
    int incl = v[N - 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl = op(n, incl);
    }
    if (lane == 31)
        s_warp[w] = incl;
    __syncthreads();
    int base = identity, total = identity;
    for (int k = 0; k < kT / 32; ++k) {
        if (k < w)
            base = op(base, s_warp[k]);
        total = op(total, s_warp[k]);
    }
    __syncthreads();
    int prev = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0)
        prev = identity;
    const int ex = op(base, prev);
#pragma unroll
    for (int k = 0; k < N; ++k)
        v[k] = op(ex, v[k]);
    return total;
}

// ------------------------------------------------------------------------------------------------ table compiler
// One CTA per tile.  FILL = false: count taps and groups; FILL = true: write the tables at the offsets in `info`.
template <bool FILL>
__global__ void __launch_bounds__(kT) k_compile_bicubic_tiles(const int* __restrict__ off_tab, const double2* __restrict__ frac_tab, int ox,
                                                             int oy, int ix, int tiles_x, int2* __restrict__ counts,
                                                             const int4* __restrict__ info, int* __restrict__ taps,
                                                             uint4* __restrict__ gmeta, double2* __restrict__ gfrac)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* s_pk = reinterpret_cast<unsigned long long*>(smem_raw); // [1024] (cell << 10 | point), sorted
    int* s_tk = reinterpret_cast<int*>(smem_raw + sizeof(unsigned long long) * kSortN); // [16384] tap keys
    int* s_runstart = s_tk + kMaxTapKeys;                                          // [1024]
    int* s_x = s_runstart + kSortN;                                                // [1024] groups before the element's run
    int* s_cells = s_x + kSortN;                                                   // [1024] distinct cells
    int* s_uniq = s_cells + kSortN;                                                // [kTapCap] distinct taps (FILL)
    __shared__ int s_warp[kT / 32];

    const int tile = blockIdx.x, t = threadIdx.x;
    int4 inf = make_int4(0, 0, 0, 0);
    if (FILL) {
        inf = info[tile];
        if (inf.y <= 0)
            return; // direct-fallback tile or a tile without a single valid point
    }
    const int tx = tile % tiles_x, ty = tile / tiles_x;

    // A. (cell, point) keys
#pragma unroll
    for (int k = 0; k < kSortN / kT; ++k) {
        const int p = t + k * kT;
        unsigned long long key = ~0ull;
        if (p < kTP) {
            const int gx = tx * kTX + (p & (kTX - 1)), gy = ty * kTY + (p >> 5);
            if (gx < ox && gy < oy) {
                const int o = off_tab[(long long)gy * ox + gx];
                if (o >= 0)
                    key = ((unsigned long long)(unsigned)o << 10) | (unsigned)p;
            }
        }
        s_pk[p] = key;
    }
    __syncthreads();
    // B. sort by cell (ties by point index: row-major inside the tile)
    bitonic_sort(s_pk, kSortN);

    // C. runs of equal cells -> groups of up to 4 points
    int cell[4], runstart[4], isfirst[4], contrib[4];
    bool valid[4];
    {
        const int i0 = 4 * t;
        unsigned long long prev = i0 > 0 ? s_pk[i0 - 1] : ~0ull;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned long long key = s_pk[i0 + k];
            valid[k] = key != ~0ull;
            cell[k] = (int)(key >> 10);
            const bool first = valid[k] && (i0 + k == 0 || (int)(prev >> 10) != cell[k]);
            isfirst[k] = first ? 1 : 0;
            runstart[k] = first ? i0 + k : 0;
            prev = key;
        }
    }
    block_scan<4>(runstart, OpMax(), 0, s_warp);
    int ord[4] = {isfirst[0], isfirst[1], isfirst[2], isfirst[3]};
    const int ncells = block_scan<4>(ord, OpAdd(), 0, s_warp);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = 4 * t + k;
        s_runstart[i] = runstart[k];
        if (isfirst[k])
            s_cells[ord[k] - 1] = cell[k];
        const unsigned long long next = (i + 1 < kSortN) ? s_pk[i + 1] : ~0ull;
        const bool last = valid[k] && (next == ~0ull || (int)(next >> 10) != cell[k]);
        contrib[k] = last ? (i - runstart[k] + 1 + 3) / 4 : 0;
    }
    int gincl[4] = {contrib[0], contrib[1], contrib[2], contrib[3]};
    const int ngroups = block_scan<4>(gincl, OpAdd(), 0, s_warp);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        s_x[4 * t + k] = gincl[k] - contrib[k];
    __syncthreads();

    // D. tap keys: the 16 taps of every distinct cell, sorted
    const int nkeys = ncells * 16;
    int n2 = 32;
    while (n2 < nkeys)
        n2 <<= 1;
    for (int j = t; j < n2; j += kT)
        s_tk[j] = j < nkeys ? s_cells[j >> 4] + ((j >> 2) & 3) * ix + (j & 3) : kNoKey;
    __syncthreads();
    bitonic_sort(s_tk, n2);

    // E. distinct taps
    const int per = n2 >= kT ? n2 / kT : 1;
    const int j0 = t * per;
    int cnt[1] = {0};
    if (j0 < n2) {
        for (int k = 0; k < per; ++k) {
            const int v = s_tk[j0 + k];
            cnt[0] += (v != kNoKey) && (j0 + k == 0 || s_tk[j0 + k - 1] != v);
        }
    }
    const int mine = cnt[0];
    const int ntaps = block_scan<1>(cnt, OpAdd(), 0, s_warp);
    if (!FILL) {
        if (t == 0)
            counts[tile] = make_int2(ntaps > kTapCap ? -1 : ntaps, ntaps > kTapCap ? 0 : ngroups);
        return;
    }
    {
        int pos = cnt[0] - mine;
        if (j0 < n2) {
            for (int k = 0; k < per; ++k) {
                const int v = s_tk[j0 + k];
                if ((v != kNoKey) && (j0 + k == 0 || s_tk[j0 + k - 1] != v)) {
                    s_uniq[pos] = v;
                    taps[(size_t)inf.x + pos] = v;
                    ++pos;
                }
            }
        }
    }
    __syncthreads();

    // F. groups: stencil rows as indices into the tap list, point slots, fractions
    auto find = [&](int key) {
        int lo = 0, hi = ntaps - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_uniq[mid] < key)
                lo = mid + 1;
            else
                hi = mid;
        }
        return lo;
    };
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!valid[k])
            continue;
        const int i = 4 * t + k;
        const int rs = s_runstart[i];
        const int within = i - rs;
        const size_t gid = (size_t)inf.z + s_x[rs] + (within >> 2);
        const int slot = within & 3;
        const int p = (int)(s_pk[i] & 1023u);
        const int gx = tx * kTX + (p & (kTX - 1)), gy = ty * kTY + (p >> 5);
        unsigned short* m16 = reinterpret_cast<unsigned short*>(gmeta + gid);
        m16[4 + slot] = (unsigned short)out_slot(p);
        gfrac[gid * 4 + slot] = frac_tab[(long long)gy * ox + gx];
        if (slot == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                m16[r] = (unsigned short)find(cell[k] + r * ix);
        }
        if (contrib[k]) { // last point of its cell: pad the group
            for (int s = slot + 1; s < 4; ++s) {
                m16[4 + s] = (unsigned short)kDump;
                gfrac[gid * 4 + s] = make_double2(0., 0.);
            }
        }
    }
    // G. order of the points inside each group.  In the gather, lane l of a warp stores point j of its group in step j; the 32
    // slots of a step fall into the 32 banks of the output tile like random numbers (2.8 wavefronts per store instruction).  The
    // order inside a group is free, so every warp of 32 consecutive groups is scheduled greedily: group after group takes the
    // rotation of its 4 points that collides least with the banks already taken in each step (1.9 wavefronts in a simulation of
    // config 2's tiles; all 24 permutations or further sweeps gain < 3 % more).  The arithmetic per point is unchanged.
    __syncthreads(); // the groups written above (global memory, this block) are complete
    for (int w = t; w * 32 < ngroups; w += kT) {
        unsigned char occ[4][32];
        for (int p = 0; p < 4; ++p)
            for (int b = 0; b < 32; ++b)
                occ[p][b] = 0;
        const int g1 = (w * 32 + 32 < ngroups) ? w * 32 + 32 : ngroups;
        for (int gi = w * 32; gi < g1; ++gi) {
            const size_t gid = (size_t)inf.z + gi;
            unsigned short* m16 = reinterpret_cast<unsigned short*>(gmeta + gid);
            unsigned short sl[4];
            double2 fr[4];
            for (int p = 0; p < 4; ++p) {
                sl[p] = m16[4 + p];
                fr[p] = gfrac[gid * 4 + p];
            }
            int best = 0, best_cost = 1 << 30;
            for (int r = 0; r < 4; ++r) {
                int cost = 0;
                for (int p = 0; p < 4; ++p) {
                    const unsigned short v = sl[(p + r) & 3];
                    if (v != (unsigned short)kDump)
                        cost += occ[p][v & 31];
                }
                if (cost < best_cost) {
                    best_cost = cost;
                    best = r;
                }
            }
            for (int p = 0; p < 4; ++p) {
                const unsigned short v = sl[(p + best) & 3];
                m16[4 + p] = v;
                gfrac[gid * 4 + p] = fr[(p + best) & 3];
                if (v != (unsigned short)kDump)
                    ++occ[p][v & 31];
            }
        }
    }
}

constexpr size_t kCompileSmem = sizeof(unsigned long long) * kSortN + sizeof(int) * (kMaxTapKeys + 3 * kSortN + kTapCap);

// 4-byte asynchronous global -> shared copy (LDGSTS) and its group control: the fp32 mode stages its taps with these (no
// conversion is needed, so the values never pass through registers)
__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gmem_src)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_and_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ gather
// Arithmetic of the gather (template tag ARITH):
//   kExact    the reference's operation order in fp64, every multiply and add rounded separately, the accumulator re-rounded
//             to fp32 after each stencil row: bit-identical (default)
//   kContract fp64 FMA chains, one final rounding (opt-in FIMEX_B200_BICUBIC_CONTRACT=1)
//   kFp32     opt-in FIMEX_B200_BICUBIC_FP32=1: the separable weights are computed in fp64 exactly as the reference computes
//             them (interpolation.c:962-1000) and rounded to fp32 ONCE per point; taps stay fp32 (no conversion at all) and the
//             20 multiply-adds per output are fp32 FMAs.  Not bit-identical; the NaN mask is (every tap enters every product
//             chain), the values agree within 1e-5 of the largest tap of the 4x4 stencil -- north_star's bar for bicubic.  It
//             takes the kernel off the fp64 pipe (35 fp64 instructions per output at 64 lanes/clk/SM) onto the fp32 pipe.
enum { kExact = 0, kContract = 1, kFp32 = 2 };
template <int ARITH>
struct ArithTypes {
    typedef double tap; // type of a staged tap and of a weight
};
template <>
struct ArithTypes<kFp32> {
    typedef float tap;
};
// weight type held in registers per group: fp32 mode keeps every weight twice, (w, w), the second operand of Blackwell's
// packed fma.rn.f32x2 (FFMA2): two LEVELS of a tap per instruction, 10 instead of 20 FP instructions per output
template <int ARITH>
struct WeightType {
    typedef double type;
};
template <>
struct WeightType<kFp32> {
    typedef float2 type;
};
__device__ __forceinline__ void set_weight(double& w, double v) { w = v; }
__device__ __forceinline__ void set_weight(float2& w, double v) { w = make_float2((float)v, (float)v); }

template <class W>
struct GroupT {
    W wx[4][4], wy[4][4];
    int row[4]; // first tap of each stencil row, in elements from the start of a field's staging area
    int pt[4];  // output-tile slots
};

template <int S, class W>
__device__ __forceinline__ void load_group(GroupT<W>& gr, const uint4* __restrict__ gmeta, const double2* __restrict__ gfrac, size_t gid)
{
    const uint4 m = __ldg(gmeta + gid);
    gr.row[0] = (int)(m.x & 0xffffu) * S;
    gr.row[1] = (int)(m.x >> 16) * S;
    gr.row[2] = (int)(m.y & 0xffffu) * S;
    gr.row[3] = (int)(m.y >> 16) * S;
    gr.pt[0] = (int)(m.z & 0xffffu);
    gr.pt[1] = (int)(m.z >> 16);
    gr.pt[2] = (int)(m.w & 0xffffu);
    gr.pt[3] = (int)(m.w >> 16);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const double2 f = __ldg(gfrac + gid * 4 + p);
        double wx[4], wy[4];
        cubic_weights(f.x, wx); // fp64, the reference's own sums; rounded once when W is float2.  (A table of the rounded weights,
        cubic_weights(f.y, wy); // 32 bytes per point, was measured: 25.2 ms instead of 24.0 -- recomputing per chunk is faster.)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            set_weight(gr.wx[p][i], wx[i]);
            set_weight(gr.wy[p][i], wy[i]);
        }
    }
}

// NL levels of one group: 4 points x NF fields x NL levels = 4*NF*NL independent accumulator chains for the scheduler.
// EXACT: the reference's operation order, every multiply and add rounded separately and the accumulator re-rounded to fp32
// after each stencil row (35 fp64 instructions + 7 conversions per output).  !EXACT (opt-in, FIMEX_B200_BICUBIC_CONTRACT=1):
// the same sums as fp64 FMA chains with ONE final rounding to fp32 (20 fp64 instructions + 1 conversion) -- not bit-identical,
// within 1e-5 of the field's magnitude (the tolerance the north star states for interpolated floats), since every
// intermediate is at least as accurate as the reference's.
// fp64 arithmetic + rotation + copy-engine stores: 40 KB of taps + 64 KB of output tiles + 14 KB of (cos, sin) would not leave room
// for two CTAs per SM, so that one combination reads (cos, sin) from global memory instead
constexpr bool cs_in_shared(int arith, bool rot, bool tmaout)
{
    return rot && !(tmaout && arith != 2 /* kFp32 */);
}

// (cos, sin) of the rotation by output-tile slot: from the shared-memory copy of the tile's values, or -- where that copy would
// push two CTAs per SM over the shared-memory limit (exact arithmetic + copy-engine stores) -- from global memory through L1
// (the same 4 addresses per thread in every batch of a chunk)
struct CsSrc {
    const double2* smem;  // [kOutRow] by slot, or null
    const double2* glob;  // the (cos, sin) table of the whole target grid
    long long base;       // index of the tile's first point
    int ox;
    __device__ __forceinline__ double2 get(int slot) const
    {
        if (smem)
            return smem[slot];
        if (slot >= kTP)
            return make_double2(1., 0.); // padding slot
        const int p = out_slot(slot); // out_slot is an involution: slot -> tile point y * 32 + x
        return __ldg(glob + base + (long long)(p >> 5) * ox + (p & 31));
    }
};

// TileConv: what is applied to a result on its way into the output tile -- nothing (the store phase converts), or, when the tile
// leaves through the copy engine as it is, the float form of interpolationArray2Data (NaN -> fill, -0 -> +0)
struct TileRaw {
    __device__ __forceinline__ float operator()(float v) const { return v; }
};
template <int NF, bool ROT, int S, int NL, int ARITH, int OR, class TileConv>
__device__ __forceinline__ void compute_levels(const GroupT<double>& gr, const double* __restrict__ st, int field_stride,
                                               float* __restrict__ s_out, int out_field_stride, const CsSrc& s_cs, const TileConv& tc)
{
    constexpr bool EXACT = ARITH == kExact;
    float a[NL][NF][4];
    if (EXACT) {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const double* sf = st + f * field_stride;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double* rp = sf + gr.row[r];
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                    const double v0 = rp[l], v1 = rp[S + l], v2 = rp[2 * S + l], v3 = rp[3 * S + l];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const double row = bicubic_row(gr.wx[p], v0, v1, v2, v3);
                        a[l][f][p] = (r == 0) ? bicubic_acc<true>(0.f, row, gr.wy[p][0]) : bicubic_acc<false>(a[l][f][p], row, gr.wy[p][r]);
                    }
                }
            }
        }
    } else {
        double d[NL][NF][4];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const double* sf = st + f * field_stride;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double* rp = sf + gr.row[r];
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                    const double v0 = rp[l], v1 = rp[S + l], v2 = rp[2 * S + l], v3 = rp[3 * S + l];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const double row = fma(gr.wx[p][3], v3, fma(gr.wx[p][2], v2, fma(gr.wx[p][1], v1, gr.wx[p][0] * v0)));
                        d[l][f][p] = (r == 0) ? row * gr.wy[p][0] : fma(row, gr.wy[p][r], d[l][f][p]);
                    }
                }
            }
        }
#pragma unroll
        for (int l = 0; l < NL; ++l)
#pragma unroll
            for (int f = 0; f < NF; ++f)
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    a[l][f][p] = (float)d[l][f][p];
    }
    if (ROT) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const double2 c = s_cs.get(gr.pt[p]);
#pragma unroll
            for (int l = 0; l < NL; ++l)
                rotate_uv(a[l][0][p], a[l][NF - 1][p], c.x, c.y);
        }
    }
#pragma unroll
    for (int l = 0; l < NL; ++l)
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int p = 0; p < 4; ++p)
                s_out[f * out_field_stride + l * OR + gr.pt[p]] = tc(a[l][f][p]);
}


// kFp32: fp32 taps, fp32 weights, 20 multiply-adds per output.  NL = 4 reads four levels of a tap with one 128-bit shared load
// (the staging buffer is tap-major with a 16-byte aligned stride) and does the arithmetic on level PAIRS with packed
// fma.rn.f32x2; NL = 1 (partial batches, many-tap tiles) is scalar.  The rotation is done in fp32 as well.
template <int NF, bool ROT, int S, int NL, int ARITH, int OR, class TileConv>
__device__ __forceinline__ void compute_levels(const GroupT<float2>& gr, const float* __restrict__ st, int field_stride, float* __restrict__ s_out,
                                               int out_field_stride, const CsSrc& s_cs, const TileConv& tc)
{
    static_assert(NL == 1 || NL == 4, "one level or one 128-bit load of four");
    float a[NL][NF][4];
    if (NL == 4) {
        float2 acc[2][NF][4]; // [level pair][field][point]
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const float* sf = st + f * field_stride;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float* rp = sf + gr.row[r];
                float2 v[4][2];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 q = *reinterpret_cast<const float4*>(rp + c * S);
                    v[c][0] = make_float2(q.x, q.y);
                    v[c][1] = make_float2(q.z, q.w);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float2 row = __ffma2_rn(gr.wx[p][3], v[3][h],
                                                      __ffma2_rn(gr.wx[p][2], v[2][h], __ffma2_rn(gr.wx[p][1], v[1][h], __fmul2_rn(gr.wx[p][0], v[0][h]))));
                        acc[h][f][p] = (r == 0) ? __fmul2_rn(row, gr.wy[p][0]) : __ffma2_rn(row, gr.wy[p][r], acc[h][f][p]);
                    }
                }
            }
        }
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                a[0][f][p] = acc[0][f][p].x;
                a[1 % NL][f][p] = acc[0][f][p].y;
                a[2 % NL][f][p] = acc[1][f][p].x;
                a[3 % NL][f][p] = acc[1][f][p].y;
            }
    } else {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const float* sf = st + f * field_stride;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float* rp = sf + gr.row[r];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float row = fmaf(gr.wx[p][3].x, rp[3 * S], fmaf(gr.wx[p][2].x, rp[2 * S], fmaf(gr.wx[p][1].x, rp[S], gr.wx[p][0].x * rp[0])));
                    a[0][f][p] = (r == 0) ? row * gr.wy[p][0].x : fmaf(row, gr.wy[p][r].x, a[0][f][p]);
                }
            }
        }
    }
    if (ROT) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const double2 c = s_cs.get(gr.pt[p]);
            const float cc = (float)c.x, ss = (float)c.y;
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                const float u = a[l][0][p], w = a[l][NF - 1][p];
                a[l][0][p] = fmaf(u, cc, -(w * ss));
                a[l][NF - 1][p] = fmaf(u, ss, w * cc);
            }
        }
    }
#pragma unroll
    for (int l = 0; l < NL; ++l)
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int p = 0; p < 4; ++p)
                s_out[f * out_field_stride + l * OR + gr.pt[p]] = tc(a[l][f][p]);
}

// Tiles whose taps do not fit the staging buffers: every point reads its 16 taps from global memory (the arithmetic of
// the direct kernel, gather_kernels.cu); lane = x, so a warp still writes 128 contiguous bytes per level.
// bicubic_eval with mifi_bad2nanf applied to the 16 taps
__device__ __forceinline__ float bicubic_eval_bad(const float* __restrict__ s, int ix, const double (&wx)[4], const double (&wy)[4], float bad)
{
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float* p = s + r * ix;
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            v[c] = __ldg(p + c);
            if (v[c] == bad)
                v[c] = undef_f();
        }
        const double row = bicubic_row(wx, (double)v[0], (double)v[1], (double)v[2], (double)v[3]);
        acc = (r == 0) ? bicubic_acc<true>(acc, row, wy[0]) : bicubic_acc<false>(acc, row, wy[r]);
    }
    return acc;
}

template <int NF, bool ROT, class Out>
__device__ void direct_tile(const GatherGeom& g, int tx, int ty, long long z0, long long z1, const int* __restrict__ off_tab,
                            const double2* __restrict__ frac_tab, const double2* __restrict__ cs, const float* __restrict__ in0,
                            const float* __restrict__ in1, typename Out::type* __restrict__ out0, typename Out::type* __restrict__ out1,
                            const Out& conv, bool fill_in, float bad0, float bad1)
{
    for (int p = threadIdx.x; p < kTP; p += kT) {
        const int gx = tx * kTX + (p & (kTX - 1)), gy = ty * kTY + (p >> 5);
        if (gx >= g.ox || gy >= g.oy)
            continue;
        const long long q = (long long)gy * g.ox + gx;
        const int off = __ldg(off_tab + q);
        typename Out::type* o0 = out0 + z0 * g.out_level + q;
        typename Out::type* o1 = (NF == 2) ? out1 + z0 * g.out_level + q : nullptr;
        if (off < 0) { // :1022-1026
            for (long long z = z0; z < z1; ++z) {
                __stcs(o0, conv(undef_f()));
                o0 += g.out_level;
                if (NF == 2) {
                    __stcs(o1, conv(undef_f()));
                    o1 += g.out_level;
                }
            }
            continue;
        }
        const double2 f = __ldg(frac_tab + q);
        double wx[4], wy[4];
        cubic_weights(f.x, wx);
        cubic_weights(f.y, wy);
        double2 rot = make_double2(1., 0.);
        if (ROT)
            rot = __ldg(cs + q);
        const float* p0 = in0 + z0 * g.in_level + off;
        const float* p1 = (NF == 2) ? in1 + z0 * g.in_level + off : nullptr;
        for (long long z = z0; z < z1; ++z) {
            float a = fill_in ? bicubic_eval_bad(p0, g.ix, wx, wy, bad0) : bicubic_eval(p0, g.ix, wx, wy);
            float b = 0.f;
            if (NF == 2)
                b = fill_in ? bicubic_eval_bad(p1, g.ix, wx, wy, bad1) : bicubic_eval(p1, g.ix, wx, wy);
            if (ROT)
                rotate_uv(a, b, rot.x, rot.y);
            __stcs(o0, conv(a));
            p0 += g.in_level;
            o0 += g.out_level;
            if (NF == 2) {
                __stcs(o1, conv(b));
                p1 += g.in_level;
                o1 += g.out_level;
            }
        }
    }
}

// FAST: at most 256 taps; 8/NF levels per batch; warp w stages (field, level) row w of the batch and later stores
//       (field, level) row w of the output tile.
// else: up to 2048/NF taps, one level per batch, every thread stages 8/NF taps per field.
// copy-engine helpers for the output tile (TMAOUT): one cp.async.bulk.tensor.3d (UTMASTG) per (field, level) row of the tile
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* map, const void* smem, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
                 "r"((unsigned)__cvta_generic_to_shared(smem)), "r"(x), "r"(y), "r"(z)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_wait_read_all()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_shared()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// TMAOUT: plain float output whose rows are 16-byte aligned: the finished (field, level) rows of the output tile are stored by
// the copy engine (tensor map with the 128-byte swizzle, which is exactly out_slot(); partial tiles are clipped by the map) while
// the warps go on with the next batch -- no read-back, no store instructions, no edge cases in the store phase.
template <int NF, bool ROT, bool FAST, class Out, int ARITH, bool TMAOUT>
__device__ __forceinline__ void staged_tile(const GatherGeom& g, int tx, int ty, long long z0, long long z1, int4 inf,
                                            const int* __restrict__ taps, const uint4* __restrict__ gmeta, const double2* __restrict__ gfrac,
                                            const double2* __restrict__ cs, const float* __restrict__ in0, const float* __restrict__ in1,
                                            typename Out::type* __restrict__ out0, typename Out::type* __restrict__ out1, bool vec_ok,
                                            typename ArithTypes<ARITH>::tap* s_stage, float* s_out, double2* s_cs, const Out& conv,
                                            bool fill_in, float bad0, float bad1, const CUtensorMap* map0, const CUtensorMap* map1)
{
    typedef typename Out::type OutT;
    typedef typename ArithTypes<ARITH>::tap TapT;
    constexpr int L = FAST ? 8 / NF : 1;
    constexpr int OR = TMAOUT ? kOutRowTma : kOutRow; // floats per (field, level) row of the output tile
    // tap stride of the tap-major staging buffer.  fp64: odd (in doubles), so that the distinct taps of a warp land in distinct
    // banks.  fp32: a multiple of 4 floats (128-bit loads of four levels) whose first eight multiples fall into eight different
    // 16-byte bank groups: 12 for 8 levels, 4 for 4 levels.
    constexpr int S = (L == 1) ? 1 : (ARITH == kFp32 ? (L == 8 ? 12 : 4) : L + 1);
    constexpr int NLV = (ARITH == kFp32) ? 4 : kNL; // levels per inner step of a full batch
    constexpr int NREG = 8;
    static_assert(NF * L <= kOutRows, "output tile");
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int ntaps = inf.y, ngroups = inf.w;
    const int field_stride = ntaps * S;
    const int rounds = (ngroups + kT - 1) / kT;

    // output tiles: NaN everywhere; points outside the 4x4 support are in no group and stay NaN (:1022-1026)
    // with TMAOUT the tile holds the FINAL values (Out is float-valued there: StorePlain or StoreAs<float>), otherwise raw results
    // that the store phase converts
    typename std::conditional<TMAOUT, Out, TileRaw>::type tile_conv = [&]() {
        if constexpr (TMAOUT)
            return conv;
        else
            return TileRaw();
    }();
    for (int i = t; i < 2 * kOutRows * OR; i += kT)
        s_out[i] = tile_conv(undef_f());
    CsSrc cs_src;
    cs_src.smem = s_cs;
    cs_src.glob = cs;
    cs_src.base = (long long)(ty * kTY) * g.ox + (long long)tx * kTX;
    cs_src.ox = g.ox;
    if (ROT && s_cs != nullptr) {
        for (int p = t; p < kOutRow; p += kT) {
            double2 c = make_double2(1., 0.);
            int slot = p;
            if (p < kTP) {
                const int gx = tx * kTX + (p & (kTX - 1)), gy = ty * kTY + (p >> 5);
                if (gx < g.ox && gy < g.oy)
                    c = __ldg(cs + (long long)gy * g.ox + gx);
                slot = out_slot(p);
            }
            s_cs[slot] = c;
        }
    }
    GroupT<typename WeightType<ARITH>::type> gr;
    if (rounds == 1 && t < ngroups)
        load_group<S>(gr, gmeta, gfrac, (size_t)inf.z + t);

    // staging: which (field, level-in-batch, tap) this thread's register j carries
    const int my_f = FAST ? warp / L : 0, my_zi = FAST ? warp % L : 0; // FAST: one (field, level) row per warp
    auto tap_of = [&](int j) { return FAST ? lane + 32 * j : t + kT * (j % (NREG / NF)); };
    auto field_of = [&](int j) { return FAST ? my_f : j / (NREG / NF); };
    float regs[NREG];
    auto load = [&](long long z) {
        if (z + my_zi >= z1)
            return;
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
            const int r = tap_of(j);
            if (r < ntaps) {
                const float* base = (field_of(j) == 0 ? in0 : in1) + (z + my_zi) * g.in_level;
                regs[j] = __ldg(base + __ldg(taps + inf.x + r));
            }
        }
    };
    auto park = [&](int buf) { // the ONE fp32 -> fp64 conversion of each tap value (none in fp32 mode)
        TapT* dst = s_stage + buf * stage_elems(ARITH == kFp32) + my_zi;
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
            const int r = tap_of(j);
            if (r < ntaps) {
                float v = regs[j];
                if (fill_in && v == (field_of(j) == 0 ? bad0 : bad1)) // mifi_bad2nanf, once per tap
                    v = undef_f();
                dst[field_of(j) * field_stride + r * S] = (TapT)v;
            }
        }
    };

    // fp32 mode: the same (field, level, tap) assignment, copied global -> shared asynchronously; batch b+1 is in flight while
    // batch b is consumed.  mifi_bad2nanf is patched in by the thread that copied the value, after the copies have landed.
    // (the source offsets of this thread's taps are the same for every batch: read once, not behind every copy)
    int tap_off[NREG];
    if constexpr (ARITH == kFp32) {
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
            const int r = tap_of(j);
            tap_off[j] = r < ntaps ? __ldg(taps + inf.x + r) : -1;
        }
    }
    auto stage_async = [&](int buf, long long z) {
        if constexpr (ARITH == kFp32) {
            if (z + my_zi < z1) {
                float* dst = reinterpret_cast<float*>(s_stage) + buf * stage_elems(true) + my_zi;
#pragma unroll
                for (int j = 0; j < NREG; ++j) {
                    if (tap_off[j] >= 0) {
                        const float* base = (field_of(j) == 0 ? in0 : in1) + (z + my_zi) * g.in_level;
                        cp_async_f32(dst + field_of(j) * field_stride + tap_of(j) * S, base + tap_off[j]);
                    }
                }
            }
        }
    };
    auto patch_async = [&](int buf, long long z) {
        if constexpr (ARITH == kFp32) {
            if (fill_in && z + my_zi < z1) {
                float* dst = reinterpret_cast<float*>(s_stage) + buf * stage_elems(true) + my_zi;
#pragma unroll
                for (int j = 0; j < NREG; ++j) {
                    const int r = tap_of(j);
                    if (r < ntaps) {
                        float* p = dst + field_of(j) * field_stride + r * S;
                        if (*p == (field_of(j) == 0 ? bad0 : bad1))
                            *p = undef_f();
                    }
                }
            }
        }
    };

    // store phase: warp w writes (field, level) row w of the finished output tile, 7 x 128-bit per lane.  Lane l owns
    // tile rows (l >> 3) + 4k; the column group alternates between two values with the parity of k (out_slot()).
    const int st_f = warp / L, st_zi = warp % L;
    OutT* const st_out = (NF == 2 && st_f == 1) ? out1 : out0;
    const int ly0 = lane >> 3;
    const int gx_even = tx * kTX + (((lane & 7) ^ ly0) << 2), gx_odd = tx * kTX + (((lane & 7) ^ (ly0 + 4)) << 2);
    const int gy0 = ty * kTY + ly0;
    const bool whole = vec_ok && (tx + 1) * kTX <= g.ox && (ty + 1) * kTY <= g.oy; // tile entirely inside the grid
    const long long st_row0 = (long long)gy0 * g.ox;
    auto store_out = [&](const float* tile, long long z, int nb) {
        if (warp >= NF * L || st_zi >= nb)
            return;
        const float* src = tile + (st_f * L + st_zi) * OR + lane * 4;
        OutT* lvl = st_out + (z + st_zi) * g.out_level + st_row0;
        if (whole) {
#pragma unroll
            for (int k = 0; k < kTP / 128; ++k) {
                const float4 v = *reinterpret_cast<const float4*>(src + 128 * k);
                store_vec4<OutT>(lvl + (long long)(4 * k) * g.ox + ((k & 1) ? gx_odd : gx_even), conv(v.x), conv(v.y), conv(v.z), conv(v.w));
            }
            return;
        }
#pragma unroll
        for (int k = 0; k < kTP / 128; ++k) {
            const int gx = (k & 1) ? gx_odd : gx_even;
            if (gy0 + 4 * k >= g.oy || gx >= g.ox)
                continue;
            const float4 v = *reinterpret_cast<const float4*>(src + 128 * k);
            OutT* dst = lvl + (long long)(4 * k) * g.ox + gx;
            if (vec_ok && gx + 3 < g.ox) {
                store_vec4<OutT>(dst, conv(v.x), conv(v.y), conv(v.z), conv(v.w));
            } else {
                __stcs(dst, conv(v.x));
                if (gx + 1 < g.ox)
                    __stcs(dst + 1, conv(v.y));
                if (gx + 2 < g.ox)
                    __stcs(dst + 2, conv(v.z));
                if (gx + 3 < g.ox)
                    __stcs(dst + 3, conv(v.w));
            }
        }
    };

    if constexpr (ARITH == kFp32) {
        stage_async(0, z0);
        cp_async_commit_and_wait_all();
        patch_async(0, z0);
    } else {
        load(z0);
        park(0);
        if (z0 + L < z1)
            load(z0 + L);
    }
    __syncthreads();
    int buf = 0;
    for (long long z = z0; z < z1; z += L, buf ^= 1) {
        const int nb = (int)((z1 - z) < L ? (z1 - z) : L);
        const TapT* st = s_stage + buf * stage_elems(ARITH == kFp32);
        float* tile = s_out + buf * (kOutRows * OR);
        if constexpr (ARITH == kFp32) { // the next batch lands in the other buffer while this one is consumed
            if (z + L < z1)
                stage_async(buf ^ 1, z + L);
            cp_async_commit();
        }
        for (int rd = 0; rd < rounds; ++rd) {
            const int gi = rd * kT + t;
            if (gi < ngroups) {
                if (rounds > 1)
                    load_group<S>(gr, gmeta, gfrac, (size_t)inf.z + gi);
                if (L > 1 && nb == L) { // full batch: branch-free, NLV levels per step
#pragma unroll kBicUnroll
                    for (int zi = 0; zi < L; zi += NLV)
                        compute_levels<NF, ROT, S, (L > 1 ? NLV : 1), ARITH, OR>(gr, st + zi, field_stride, tile + zi * OR, L * OR, cs_src, tile_conv);
                } else {
#pragma unroll 1
                    for (int zi = 0; zi < nb; ++zi)
                        compute_levels<NF, ROT, S, 1, ARITH, OR>(gr, st + zi, field_stride, tile + zi * OR, L * OR, cs_src, tile_conv);
                }
            }
        }
        if constexpr (ARITH == kFp32) {
            cp_async_wait_all();
            if (z + L < z1)
                patch_async(buf ^ 1, z + L);
        } else if (z + L < z1) {
            park(buf ^ 1);
            if (z + 2 * L < z1)
                load(z + 2 * L);
        }
        if constexpr (TMAOUT) {
            fence_async_shared(); // this thread's tile writes, for the copy engine
            if (lane == 0)
                tma_wait_read_all(); // the store issued after the previous barrier has finished reading the OTHER tile, which the
                                     // next iteration overwrites
        }
        __syncthreads(); // this batch's output tile is complete and the next batch is parked; the other output tile (being
                         // stored by slower warps) is not written before the next barrier
        if constexpr (TMAOUT) {
            if (lane == 0 && warp < NF * L && st_zi < nb)
                tma_store_tile(st_f == 0 ? map0 : map1, tile + (st_f * L + st_zi) * OR, tx * kTX, ty * kTY, (int)(z + st_zi));
        } else {
            store_out(tile, z, nb);
        }
    }
    if constexpr (TMAOUT) {
        if (lane == 0)
            tma_wait_read_all(); // shared memory must outlive the copies that read it
    }
}

template <int NF, bool ROT, class Out, int ARITH, bool TMAOUT>
__global__ void __launch_bounds__(kT, 2) k_gather_bicubic_staged(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                                                                 GatherGeom g, int tiles_x, const int4* __restrict__ info,
                                                                 const int* __restrict__ taps, const uint4* __restrict__ gmeta,
                                                                 const double2* __restrict__ gfrac, const int* __restrict__ off_tab,
                                                                 const double2* __restrict__ frac_tab, const double2* __restrict__ cs,
                                                                 const float* __restrict__ in0, const float* __restrict__ in1,
                                                                 typename Out::type* __restrict__ out0, typename Out::type* __restrict__ out1,
                                                                 int vec_ok, long long per, Out conv, int fill_in, float bad0, float bad1)
{
    typedef typename ArithTypes<ARITH>::tap TapT;
    constexpr int OR = TMAOUT ? kOutRowTma : kOutRow;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TapT* s_stage = reinterpret_cast<TapT*>(smem_raw);                                         // [2][stage_elems] taps (fp64, or fp32 in fp32 mode)
    float* s_out = reinterpret_cast<float*>(smem_raw + sizeof(TapT) * 2 * stage_elems(ARITH == kFp32)); // [2][8][OR]; 1024-byte aligned (TMAOUT)
    double2* s_cs = cs_in_shared(ARITH, ROT, TMAOUT) ? reinterpret_cast<double2*>(s_out + 2 * kOutRows * OR) : nullptr; // [kOutRow] (ROT)
    const int tile = blockIdx.x;
    const int4 inf = __ldg(info + tile);
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const long long z0 = (long long)blockIdx.y * per;
    const long long z1 = z0 + per < g.nz ? z0 + per : g.nz;
    if (z0 >= z1)
        return;
    if (inf.y < 0 || NF * inf.y > kStageElems)
        direct_tile<NF, ROT, Out>(g, tx, ty, z0, z1, off_tab, frac_tab, cs, in0, in1, out0, out1, conv, fill_in != 0, bad0, bad1);
    else if (inf.y <= kFastTaps)
        staged_tile<NF, ROT, true, Out, ARITH, TMAOUT>(g, tx, ty, z0, z1, inf, taps, gmeta, gfrac, cs, in0, in1, out0, out1, vec_ok != 0, s_stage,
                                                       s_out, s_cs, conv, fill_in != 0, bad0, bad1, &map0, &map1);
    else
        staged_tile<NF, ROT, false, Out, ARITH, TMAOUT>(g, tx, ty, z0, z1, inf, taps, gmeta, gfrac, cs, in0, in1, out0, out1, vec_ok != 0, s_stage,
                                                        s_out, s_cs, conv, fill_in != 0, bad0, bad1, &map0, &map1);
}

constexpr size_t gather_smem(int arith, bool rot, bool tmaout)
{
    return (arith == kFp32 ? sizeof(float) : sizeof(double)) * 2 * stage_elems(arith == kFp32) +
           sizeof(float) * 2 * kOutRows * (tmaout ? kOutRowTma : kOutRow) + (cs_in_shared(arith, rot, tmaout) ? sizeof(double2) * kOutRow : 0);
}
// two CTAs per SM must still fit (228 KB per SM, 1 KB reserved per CTA)
constexpr bool tma_store_fits(int arith, bool rot)
{
    return 2 * (gather_smem(arith, rot, true) + 1024) <= 228 * 1024;
}

} // namespace

// ------------------------------------------------------------------------------------------------- host side
void bicubic_tiles_free(BicubicTiles* bt)
{
    if (bt->d_info)
        cudaFree(bt->d_info);
    if (bt->d_taps)
        cudaFree(bt->d_taps);
    if (bt->d_gmeta)
        cudaFree(bt->d_gmeta);
    if (bt->d_gfrac)
        cudaFree(bt->d_gfrac);
    *bt = BicubicTiles();
}

bool bicubic_tiles_supported(int ix, int iy, int ox, int oy)
{
    const long long in_level = (long long)ix * iy;
    const long long tiles = (long long)((ox + kTX - 1) / kTX) * ((oy + kTY - 1) / kTY);
    // cell offsets are packed above 10 bits of point index in a 64-bit key and taps are ints: any int-sized level works
    return in_level > 0 && in_level < 2147483647LL - 4ll * ix && ox > 0 && oy > 0 && tiles < 2147483647LL;
}

// synchronises (setup path)
int bicubic_tiles_build(const int* d_off, const double2* d_frac, int ix, int iy, int ox, int oy, BicubicTiles* bt, cudaStream_t st)
{
    bicubic_tiles_free(bt);
    const int tiles_x = (ox + kTX - 1) / kTX, tiles_y = (oy + kTY - 1) / kTY;
    const size_t tiles = (size_t)tiles_x * tiles_y;
    FB_CUDA_CHECK(cudaFuncSetAttribute(k_compile_bicubic_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCompileSmem));
    FB_CUDA_CHECK(cudaFuncSetAttribute(k_compile_bicubic_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCompileSmem));
    int2* d_counts = nullptr;
    FB_CUDA_CHECK(cudaMalloc(&d_counts, sizeof(int2) * tiles));
    k_compile_bicubic_tiles<false><<<(unsigned)tiles, kT, kCompileSmem, st>>>(d_off, d_frac, ox, oy, ix, tiles_x, d_counts, nullptr, nullptr,
                                                                             nullptr, nullptr);
    count_launch();
    std::vector<int2> counts(tiles);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(counts.data(), d_counts, sizeof(int2) * tiles, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(st);
    cudaFree(d_counts);
    FB_CUDA_CHECK(e);
    std::vector<int4> info(tiles);
    long long ntaps = 0, ngroups = 0;
    int ndirect = 0;
    for (size_t i = 0; i < tiles; ++i) {
        FB_REQUIRE(ntaps < 2147483647LL - kTapCap && ngroups < 2147483647LL - kTP, "bicubic tile table too large");
        info[i] = make_int4((int)ntaps, counts[i].x, (int)ngroups, counts[i].y);
        if (counts[i].x < 0) {
            ++ndirect;
            continue;
        }
        ntaps += counts[i].x;
        ngroups += counts[i].y;
    }
    bt->tiles_x = tiles_x;
    bt->tiles_y = tiles_y;
    bt->n_taps = ntaps;
    bt->n_groups = ngroups;
    bt->n_direct = ndirect;
    FB_CUDA_CHECK(cudaMalloc(&bt->d_info, sizeof(int4) * tiles));
    FB_CUDA_CHECK(cudaMalloc(&bt->d_taps, sizeof(int) * (size_t)(ntaps > 0 ? ntaps : 1)));
    FB_CUDA_CHECK(cudaMalloc(&bt->d_gmeta, sizeof(uint4) * (size_t)(ngroups > 0 ? ngroups : 1)));
    FB_CUDA_CHECK(cudaMalloc(&bt->d_gfrac, sizeof(double2) * 4 * (size_t)(ngroups > 0 ? ngroups : 1)));
    FB_CUDA_CHECK(cudaMemcpyAsync(bt->d_info, info.data(), sizeof(int4) * tiles, cudaMemcpyHostToDevice, st));
    k_compile_bicubic_tiles<true><<<(unsigned)tiles, kT, kCompileSmem, st>>>(d_off, d_frac, ox, oy, ix, tiles_x, nullptr, bt->d_info, bt->d_taps,
                                                                            bt->d_gmeta, bt->d_gfrac);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    FB_CUDA_CHECK(cudaStreamSynchronize(st)); // `info` (host) is read by the copy above
    return FB_OK;
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

// [z][y][x] float output as a 3-D tensor of 32 x 28 x 1 boxes with the 128-byte swizzle (== out_slot())
bool output_tile_map(void* d_out, const GatherGeom& g, CUtensorMap* map)
{
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc || !d_out)
        return false;
    const cuuint64_t dims[3] = {(cuuint64_t)g.ox, (cuuint64_t)g.oy, (cuuint64_t)g.nz};
    const cuuint64_t strides[2] = {(cuuint64_t)g.ox * sizeof(float), (cuuint64_t)g.out_level * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)kTX, (cuuint32_t)kTY, 1u};
    const cuuint32_t es[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NF, bool ROT, int ARITH, class Out, bool TMAOUT>
int launch_bic_as(dim3 grid, const CUtensorMap& map0, const CUtensorMap& map1, const GatherGeom& g, const BicubicTiles& bt, const int* d_off,
                  const double2* d_frac, const double2* d_cs, const float* d_in0, const float* d_in1, void* d_out0, void* d_out1, int vec_ok,
                  long long per, Out conv, const SliceConv& sc, cudaStream_t st)
{
    typedef typename Out::type T;
    constexpr size_t smem = gather_smem(ARITH, ROT, TMAOUT);
    auto kernel = k_gather_bicubic_staged<NF, ROT, Out, ARITH, TMAOUT>;
    FB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, kT, smem, st>>>(map0, map1, g, bt.tiles_x, bt.d_info, bt.d_taps, bt.d_gmeta, bt.d_gfrac, d_off, d_frac, d_cs, d_in0, d_in1,
                                   static_cast<T*>(d_out0), static_cast<T*>(d_out1), vec_ok, per, conv, sc.fill_in ? 1 : 0, sc.bad_in[0],
                                   sc.bad_in[1]);
    return FB_OK;
}

// FIMEX_B200_BICUBIC_TMA=0: per-thread stores for every output (A/B; the copy-engine store is the default where it applies)
bool bicubic_tma_wanted()
{
    const char* env = std::getenv("FIMEX_B200_BICUBIC_TMA");
    return !(env && env[0] == '0');
}

template <int NF, bool ROT, int ARITH, class Out>
int launch_bic(dim3 grid, const GatherGeom& g, const BicubicTiles& bt, const int* d_off, const double2* d_frac, const double2* d_cs,
               const float* d_in0, const float* d_in1, void* d_out0, void* d_out1, int vec_ok, long long per, Out conv, const SliceConv& sc,
               cudaStream_t st)
{
    CUtensorMap map0, map1;
    std::memset(&map0, 0, sizeof(map0));
    std::memset(&map1, 0, sizeof(map1));
    // float output (plain, or with NaN -> fill), rows and levels 16-byte aligned, two CTAs per SM still fit: the output tile leaves
    // through the copy engine
    if constexpr ((std::is_same<Out, StorePlain>::value || std::is_same<Out, StoreAs<float>>::value) && tma_store_fits(ARITH, ROT)) {
        if (vec_ok && g.nz < 2147483647LL && bicubic_tma_wanted() && output_tile_map(d_out0, g, &map0) &&
            (NF == 1 || output_tile_map(d_out1, g, &map1)))
            return launch_bic_as<NF, ROT, ARITH, Out, true>(grid, map0, map1, g, bt, d_off, d_frac, d_cs, d_in0, d_in1, d_out0, d_out1, vec_ok, per,
                                                            conv, sc, st);
    }
    return launch_bic_as<NF, ROT, ARITH, Out, false>(grid, map0, map1, g, bt, d_off, d_frac, d_cs, d_in0, d_in1, d_out0, d_out1, vec_ok, per, conv,
                                                     sc, st);
}

// plain float output in one of the three arithmetic modes
template <int NF, bool ROT>
int launch_bic_plain(int arith, dim3 grid, const GatherGeom& g, const BicubicTiles& bt, const int* d_off, const double2* d_frac,
                     const double2* d_cs, const float* d_in0, const float* d_in1, void* d_out0, void* d_out1, int vec_ok, long long per,
                     const SliceConv& sc, cudaStream_t st)
{
    switch (arith) {
    case kFp32:
        return launch_bic<NF, ROT, kFp32>(grid, g, bt, d_off, d_frac, d_cs, d_in0, d_in1, d_out0, d_out1, vec_ok, per, StorePlain(), sc, st);
    case kContract:
        return launch_bic<NF, ROT, kContract>(grid, g, bt, d_off, d_frac, d_cs, d_in0, d_in1, d_out0, d_out1, vec_ok, per, StorePlain(), sc, st);
    default:
        return launch_bic<NF, ROT, kExact>(grid, g, bt, d_off, d_frac, d_cs, d_in0, d_in1, d_out0, d_out1, vec_ok, per, StorePlain(), sc, st);
    }
}

// FIMEX_B200_BICUBIC_FP32=1: fp32 weights and FMAs; FIMEX_B200_BICUBIC_CONTRACT=1: fp64 FMA chains with one final rounding
// (see the ARITH tags above).  Read at every launch, default off: the shipped behaviour is bit-identical to the reference.
int bicubic_arith()
{
    const char* env = std::getenv("FIMEX_B200_BICUBIC_FP32");
    if (env && env[0] == '1')
        return kFp32;
    env = std::getenv("FIMEX_B200_BICUBIC_CONTRACT");
    return (env && env[0] == '1') ? kContract : kExact;
}
} // namespace

// scalar field (d_in1 == d_out1 == nullptr) or the two components of a vector, rotated when d_cs != nullptr.  The scalar
// form converts to sc.out_type while storing (staged_store_supports()); the vector form writes plain floats.
int launch_gather_bicubic_staged(const GatherGeom& g, const BicubicTiles& bt, const int* d_off, const double2* d_frac, const double2* d_cs,
                                 const float* d_in0, const float* d_in1, void* d_out0, void* d_out1, const SliceConv& sc, cudaStream_t st)
{
    if (g.out_level == 0 || g.nz == 0)
        return FB_OK;
    const unsigned tiles = (unsigned)bt.tiles_x * (unsigned)bt.tiles_y;
    // Level chunks: 128 levels where the copy engine stores the output tile (64 with per-thread stores), always a multiple of the 8-level batch so that only the last chunk has a partial batch; shorter
    // chunks only to fill the SMs of small grids.  (Round 1, per-thread stores: 64-level chunks, because CTAs that own long
    // level ranges drift apart and the set of DRAM pages being written grows -- 3288-level chunks cost 20 % more time.  With the
    // output tile stored by the copy engine, same box: fp32 mode 64 / 96 / 128 / 192 levels -> 20.24 / 19.59 / 19.35 / 19.23 ms,
    // exact 64 / 128 -> 44.42 / 43.76 ms: per-chunk set-up -- group weights, tile initialisation -- outweighs the drift now.)
    const bool two = d_in1 != nullptr;
    const size_t elem = (two || !sc.convert_out) ? sizeof(float) : type_size(sc.out_type);
    const uintptr_t align = reinterpret_cast<uintptr_t>(d_out0) | (two ? reinterpret_cast<uintptr_t>(d_out1) : 0);
    const int vec_ok = ((g.ox % 4) == 0 && (align & (4 * elem - 1)) == 0) ? 1 : 0;
    const bool engine_stores = vec_ok && (!sc.convert_out || sc.out_type == FB_T_FLOAT) && bicubic_tma_wanted() && tensor_map_encoder() != nullptr;
    long long per = engine_stores ? 128 : 64;
    const long long want = 4ll * sm_count();
    while (per > 8 && tiles * ((g.nz + per - 1) / per) < want)
        per -= 8;
    long long gy = (g.nz + per - 1) / per;
    if (const char* env = std::getenv("FIMEX_B200_ZCHUNK"))
        if (std::atoll(env) > 0)
            per = std::atoll(env), gy = (g.nz + per - 1) / per;
    if (gy > 65535) {
        gy = 65535;
        per = (g.nz + gy - 1) / gy;
    }
    dim3 grid(tiles, (unsigned)gy);
    int rc = FB_ERROR;
    const int arith = bicubic_arith();
    if (two) {
        FB_REQUIRE(!sc.convert_out, "bicubic vector gather writes plain floats");
        rc = d_cs ? launch_bic_plain<2, true>(arith, grid, g, bt, d_off, d_frac, d_cs, d_in0, d_in1, d_out0, d_out1, vec_ok, per, sc, st)
                  : launch_bic_plain<2, false>(arith, grid, g, bt, d_off, d_frac, d_cs, d_in0, d_in1, d_out0, d_out1, vec_ok, per, sc, st);
    } else if (!sc.convert_out) {
        rc = launch_bic_plain<1, false>(arith, grid, g, bt, d_off, d_frac, d_cs, d_in0, d_in1, d_out0, d_out1, vec_ok, per, sc, st);
    } else {
        switch (sc.out_type) {
#define FB_CASE(TAG, T)                                                                                                                    \
    case TAG:                                                                                                                              \
        rc = arith == kFp32 ? launch_bic<1, false, kFp32>(grid, g, bt, d_off, d_frac, d_cs, d_in0, d_in1, d_out0, d_out1, vec_ok, per,     \
                                                          StoreAs<T>{cast_fill<T>(sc.fill_out)}, sc, st)                                   \
                            : launch_bic<1, false, kExact>(grid, g, bt, d_off, d_frac, d_cs, d_in0, d_in1, d_out0, d_out1, vec_ok, per,    \
                                                           StoreAs<T>{cast_fill<T>(sc.fill_out)}, sc, st);                                 \
        break;
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
            FB_REQUIRE(false, "bicubic staged gather: unsupported output type");
        }
    }
    if (rc != FB_OK)
        return rc;
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

} // namespace fb
