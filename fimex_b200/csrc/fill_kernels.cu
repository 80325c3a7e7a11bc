// fimex_b200/csrc/fill_kernels.cu -- the 2-D pre/post-processes of CDMInterpolator::getDataSlice on the device
// (SURVEY.md 8f rank 3): mifi_fill2d_f, mifi_creepfill2d_f, mifi_creepfillval2d_f
// (/root/reference/src/interpolation.c:1246-1376, :1378-1537; hooks src/CDMInterpolator.cc:126-159, 256, 284).
//
// Both are LEXICOGRAPHIC Gauss-Seidel sweeps: cell (x, y) of sweep n reads its left and upper neighbours as updated in the
// same sweep and its right and lower neighbours from the sweep before.  The only parallel order with bit-identical results
// is the anti-diagonal wavefront: all cells with the same x + y are independent.  One CTA owns one level (levels are
// independent, processArray_ at src/CDMInterpolator.cc:136-159 loops over them), walks the diagonals with one barrier per
// diagonal, and works in place on the row-major level: the 8 following diagonals live in the same 32-byte sectors, so the
// strided accesses hit L1.  The sequential fp64 sums of the reference (mean, mean deviation) are replayed in the reference's
// order by one thread per level from coalesced shared-memory chunks, so that the first guess and the convergence threshold
// carry the same roundings.
//
// nx, ny >= 2 is required (for ny == 1 the reference itself reads row 1, interpolation.c:1361-1364).
#include "../../include/fimex_b200.h"

#include "kernels.h"

#include <cstdlib>

namespace fb {
namespace {

constexpr int kT = 512;
constexpr int kChunk = 4096;

struct Stats {
    unsigned long long n_nan;
    double sum;
};

// NaN count and the sequential double sum of the defined values in row-major order (interpolation.c:1252-1264, :1492-1503)
// sequential sum of |v - average| over the defined values of a chunk in shared memory, continuing `dev` (interpolation.c:1296)
__device__ __forceinline__ double chunk_deviation(const float* s_buf, int m, double average, double dev)
{
    int i = 0;
    for (; i + 8 <= m; i += 8) {
        const float4 a = *reinterpret_cast<const float4*>(s_buf + i), b = *reinterpret_cast<const float4*>(s_buf + i + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (!isnan(v[k]))
                dev = __dadd_rn(dev, fabs(__dsub_rn((double)v[k], average)));
    }
    for (; i < m; ++i)
        if (!isnan(s_buf[i]))
            dev = __dadd_rn(dev, fabs(__dsub_rn((double)s_buf[i], average)));
    return dev;
}

__device__ Stats level_sum(const float* __restrict__ f, size_t n, float* s_buf)
{
    __shared__ Stats s_st;
    if (threadIdx.x == 0) {
        s_st.n_nan = 0;
        s_st.sum = 0.;
    }
    for (size_t c0 = 0; c0 < n; c0 += kChunk) {
        const int m = (int)(n - c0 < (size_t)kChunk ? n - c0 : (size_t)kChunk);
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += kT)
            s_buf[i] = f[c0 + i];
        __syncthreads();
        if (threadIdx.x == 0) { // the serial chain is one DADD per value; loads, conversions and NaN tests run ahead of it
            double sum = s_st.sum;
            unsigned long long nn = s_st.n_nan;
            int i = 0;
            for (; i + 8 <= m; i += 8) {
                const float4 a = *reinterpret_cast<const float4*>(s_buf + i), b = *reinterpret_cast<const float4*>(s_buf + i + 4);
                const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (isnan(v[k]))
                        ++nn;
                    else
                        sum = __dadd_rn(sum, (double)v[k]);
                }
            }
            for (; i < m; ++i) {
                const float v = s_buf[i];
                if (isnan(v))
                    ++nn;
                else
                    sum = __dadd_rn(sum, (double)v);
            }
            s_st.sum = sum;
            s_st.n_nan = nn;
        }
    }
    __syncthreads();
    return s_st;
}

// ------------------------------------------------------------------------------------------------------- fill2d
__global__ void __launch_bounds__(kT) k_fill2d(float* __restrict__ field, float* __restrict__ wfield, int nx, int ny, float relaxCrit,
                                               float corrEff, unsigned long long maxLoop, unsigned long long* __restrict__ n_changed)
{
    __shared__ __align__(16) float s_buf[kChunk];
    __shared__ double s_dev;
    __shared__ int s_bad;
    const size_t n = (size_t)nx * ny;
    float* f = field + blockIdx.x * n;
    float* w = wfield + blockIdx.x * n;
    const int t = threadIdx.x;
    const Stats st = level_sum(f, n, s_buf);
    if (t == 0 && n_changed)
        n_changed[blockIdx.x] = st.n_nan;
    const unsigned long long nUnchanged = n - st.n_nan;
    if (nUnchanged == 0 || st.n_nan == 0)
        return; // nothing to do (:1266-1268)
    const double average = __ddiv_rn(st.sum, (double)nUnchanged);
    // first guess, weights, and the sequential mean deviation (:1286-1302)
    if (t == 0)
        s_dev = 0.;
    for (size_t c0 = 0; c0 < n; c0 += kChunk) {
        const int m = (int)(n - c0 < (size_t)kChunk ? n - c0 : (size_t)kChunk);
        __syncthreads();
        for (int i = t; i < m; i += kT) {
            const float v = f[c0 + i];
            s_buf[i] = v;
            if (isnan(v)) {
                w[c0 + i] = 1.f;
                f[c0 + i] = (float)average;
            } else {
                w[c0 + i] = 0.f;
            }
        }
        __syncthreads();
        if (t == 0)
            s_dev = chunk_deviation(s_buf, m, average, s_dev);
    }
    __syncthreads();
    const double stddev = __ddiv_rn(s_dev, (double)nUnchanged);
    const double crit = __dmul_rn((double)relaxCrit, stddev);
    const float crtest = (float)__dmul_rn(crit, (double)corrEff); // :1332
    const int nxm1 = nx - 1, nym1 = ny - 1;
    // the variational field: interior weights times corrEff (:1314-1318)
    for (size_t i = t; i < n; i += kT) {
        const int x = (int)(i % nx), y = (int)(i / nx);
        if (x >= 1 && x < nxm1 && y >= 1 && y < nym1)
            w[i] = __fmul_rn(w[i], corrEff);
    }
    __syncthreads();
    for (unsigned long long loop = 0; loop < maxLoop; ++loop) {
        const bool test = (loop < (maxLoop - 5ull)) && (loop % 10ull == 0); // size_t arithmetic as in the reference (:1329-1330)
        if (t == 0)
            s_bad = 0;
        int bad = 0;
        // interior sweep, anti-diagonal by anti-diagonal (:1321-1327)
        for (int d = 2; d <= nxm1 - 1 + nym1 - 1; ++d) {
            const int ylo = d - (nxm1 - 1) > 1 ? d - (nxm1 - 1) : 1;
            const int yhi = d - 1 < nym1 - 1 ? d - 1 : nym1 - 1;
            for (int y = ylo + t; y <= yhi; y += kT) {
                const size_t p = (size_t)y * nx + (d - y);
                const float c = f[p];
                const float s4 = __fadd_rn(__fadd_rn(__fadd_rn(f[p + 1], f[p - 1]), f[p + nx]), f[p - nx]);
                const float e = (float)__dsub_rn(__dmul_rn((double)s4, 0.25), (double)c);
                const float wv = w[p];
                const float ew = __fmul_rn(e, wv);
                f[p] = __fadd_rn(c, ew);
                if (test && fabs((double)ew) > (double)crtest)
                    bad = 1;
            }
            __syncthreads();
        }
        if (test) { // convergence now and then (:1329-1352)
            if (bad)
                s_bad = 1;
            __syncthreads();
            if (s_bad == 0)
                return;
            __syncthreads();
        }
        // borders (:1354-1363): columns first, then rows including the corners
        for (int y = 1 + t; y < nym1; y += kT) {
            const size_t r = (size_t)y * nx;
            f[r] = __fadd_rn(f[r], __fmul_rn(__fsub_rn(f[r + 1], f[r]), w[r]));
            f[r + nxm1] = __fadd_rn(f[r + nxm1], __fmul_rn(__fsub_rn(f[r + nxm1 - 1], f[r + nxm1]), w[r + nxm1]));
        }
        __syncthreads();
        for (int x = t; x < nx; x += kT) {
            const size_t top = (size_t)x, bot = (size_t)nym1 * nx + x;
            f[top] = __fadd_rn(f[top], __fmul_rn(__fsub_rn(f[top + nx], f[top]), w[top]));
            f[bot] = __fadd_rn(f[bot], __fmul_rn(__fsub_rn(f[bot - nx], f[bot]), w[bot]));
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------- fill2d, skewed
// The same sweeps with CONSECUTIVE SWEEPS PIPELINED.  Cell (x, y) of sweep n depends on (x-1, y), (x, y-1) of sweep n and on
// (x+1, y), (x, y+1), (x, y) of sweep n-1, so every cell with the same time tau = 2n + x + y is independent: up to B sweeps run
// one behind the other, two diagonals apart, and a level needs (nx + ny + 2B) barriers per B sweeps instead of B (nx + ny).
// The 2B + 6 diagonals in flight live in a shared-memory ring (values and weights), so a cell is read from and written to
// global memory once per block of sweeps instead of once per sweep.  The one-sided border relaxation after sweep n
// (interpolation.c:1354-1363) is scheduled into the same time line (each border cell right after the interior cell it reads);
// the last sweep of a block gets its borders from a plain pass over global memory, because a block ends where the reference
// tests convergence (every 10th sweep) and returns BEFORE that sweep's border pass.
template <int NR>
__global__ void __launch_bounds__(kT) k_fill2d_skewed(float* __restrict__ field, float* __restrict__ wfield, int nx, int ny, int by_y,
                                                      int lp, float relaxCrit, float corrEff, unsigned long long maxLoop,
                                                      unsigned long long* __restrict__ n_changed)
{
    constexpr int B = NR == 32 ? 10 : 5; // sweeps in flight: 2B + 5 <= NR live diagonals
    extern __shared__ __align__(16) float s_ring[]; // [NR][lp] values, then [NR][lp] weights
    __shared__ __align__(16) float s_buf[kChunk];
    __shared__ double s_dev;
    __shared__ int s_bad;
    float* rf = s_ring;
    float* rw = s_ring + NR * lp;
    const size_t n = (size_t)nx * ny;
    float* f = field + blockIdx.x * n;
    float* w = wfield + blockIdx.x * n;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int nwarps = kT / 32;
    const Stats st = level_sum(f, n, s_buf);
    if (t == 0 && n_changed)
        n_changed[blockIdx.x] = st.n_nan;
    const unsigned long long nUnchanged = n - st.n_nan;
    if (nUnchanged == 0 || st.n_nan == 0)
        return; // nothing to do (:1266-1268)
    const double average = __ddiv_rn(st.sum, (double)nUnchanged);
    if (t == 0)
        s_dev = 0.;
    for (size_t c0 = 0; c0 < n; c0 += kChunk) { // first guess, weights, sequential mean deviation (:1286-1302)
        const int m = (int)(n - c0 < (size_t)kChunk ? n - c0 : (size_t)kChunk);
        __syncthreads();
        for (int i = t; i < m; i += kT) {
            const float v = f[c0 + i];
            s_buf[i] = v;
            const size_t p = c0 + i;
            const int x = (int)(p % nx), y = (int)(p / nx);
            const bool interior = x >= 1 && x < nx - 1 && y >= 1 && y < ny - 1;
            if (isnan(v)) {
                w[p] = interior ? __fmul_rn(1.f, corrEff) : 1.f; // :1314-1318
                f[p] = (float)average;
            } else {
                w[p] = interior ? __fmul_rn(0.f, corrEff) : 0.f;
            }
        }
        __syncthreads();
        if (t == 0)
            s_dev = chunk_deviation(s_buf, m, average, s_dev);
    }
    __syncthreads();
    const double crit = __dmul_rn((double)relaxCrit, __ddiv_rn(s_dev, (double)nUnchanged));
    const float crtest = (float)__dmul_rn(crit, (double)corrEff); // :1332
    const int dmax = nx + ny - 2;
    auto slot = [&](int x, int y) { return ((x + y) & (NR - 1)) * lp + (by_y ? y : x); };
    auto load_diag = [&](int d) {
        if (d < 0 || d > dmax)
            return;
        const int xlo = d - (ny - 1) > 0 ? d - (ny - 1) : 0, xhi = d < nx - 1 ? d : nx - 1;
        for (int x = xlo + t; x <= xhi; x += kT) {
            const int y = d - x;
            const size_t p = (size_t)y * nx + x;
            rf[slot(x, y)] = f[p];
            rw[slot(x, y)] = w[p];
        }
    };
    // the diagonal that enters the ring comes through registers, fetched two steps before it is needed: a step must not wait
    // for a global-memory round trip (kPF values per thread cover diagonals of up to kPF * 512 cells)
    constexpr int kPF = 4;
    float pf[2][kPF], pw[2][kPF];
    auto fetch = [&](int d, float (&vf)[kPF], float (&vw)[kPF]) {
        if (d < 0 || d > dmax)
            return;
        const int xlo = d - (ny - 1) > 0 ? d - (ny - 1) : 0, xhi = d < nx - 1 ? d : nx - 1;
#pragma unroll
        for (int k = 0; k < kPF; ++k) {
            const int x = xlo + t + k * kT;
            if (x <= xhi) {
                const size_t p = (size_t)(d - x) * nx + x;
                vf[k] = f[p];
                vw[k] = w[p];
            }
        }
    };
    auto commit = [&](int d, const float (&vf)[kPF], const float (&vw)[kPF]) {
        if (d < 0 || d > dmax)
            return;
        const int xlo = d - (ny - 1) > 0 ? d - (ny - 1) : 0, xhi = d < nx - 1 ? d : nx - 1;
#pragma unroll
        for (int k = 0; k < kPF; ++k) {
            const int x = xlo + t + k * kT;
            if (x <= xhi) {
                rf[slot(x, d - x)] = vf[k];
                rw[slot(x, d - x)] = vw[k];
            }
        }
    };
    auto store_diag = [&](int d) {
        if (d < 0 || d > dmax)
            return;
        const int xlo = d - (ny - 1) > 0 ? d - (ny - 1) : 0, xhi = d < nx - 1 ? d : nx - 1;
        for (int x = xlo + t; x <= xhi; x += kT)
            f[(size_t)(d - x) * nx + x] = rf[slot(x, d - x)];
    };
    // one-sided relaxation of a border cell towards its inner neighbour (:1354-1363)
    auto border = [&](int x, int y, int xi, int yi) {
        const float a = rf[slot(x, y)];
        rf[slot(x, y)] = __fadd_rn(a, __fmul_rn(__fsub_rn(rf[slot(xi, yi)], a), rw[slot(x, y)]));
    };
    unsigned long long loop = 0;
    while (loop < maxLoop) {
        // a block of sweeps: up to B, ending at the first sweep the reference tests (:1329-1330, size_t arithmetic)
        int nb = 0;
        bool test_last = false;
        while (nb < B && loop + nb < maxLoop) {
            const unsigned long long k = loop + nb;
            ++nb;
            if ((k < (maxLoop - 5ull)) && (k % 10ull == 0)) {
                test_last = true;
                break;
            }
        }
        if (t == 0)
            s_bad = 0;
        int bad = 0;
        load_diag(0);
        load_diag(1);
        fetch(2, pf[0], pw[0]);
        fetch(3, pf[1], pw[1]);
        __syncthreads();
        const int tau_end = dmax + 2 * (nb - 1) + 4;
        for (int tau = 0; tau <= tau_end; ++tau) {
            if (tau & 1) { // diagonal tau + 2 enters the ring; diagonal tau + 4 starts its way from global memory
                commit(tau + 2, pf[1], pw[1]);
                fetch(tau + 4, pf[1], pw[1]);
            } else {
                commit(tau + 2, pf[0], pw[0]);
                fetch(tau + 4, pf[0], pw[0]);
            }
            int base = 0;
            for (int j = 0; j < nb; ++j) {
                const int d = tau - 2 * j;
                const int ylo = d - (nx - 2) > 1 ? d - (nx - 2) : 1;
                const int yhi = d - 1 < ny - 2 ? d - 1 : ny - 2;
                const int cnt = yhi >= ylo ? yhi - ylo + 1 : 0;
                const int nchunk = (cnt + 31) >> 5;
                int k = (warp - base % nwarps + nwarps) % nwarps;
                for (; k < nchunk; k += nwarps) {
                    const int y = ylo + 32 * k + lane;
                    if (y <= yhi) {
                        const int x = d - y;
                        const int me = slot(x, y);
                        const float c = rf[me];
                        const float s4 = __fadd_rn(__fadd_rn(__fadd_rn(rf[slot(x + 1, y)], rf[slot(x - 1, y)]), rf[slot(x, y + 1)]),
                                                   rf[slot(x, y - 1)]);
                        const float e = (float)__dsub_rn(__dmul_rn((double)s4, 0.25), (double)c);
                        const float ew = __fmul_rn(e, rw[me]);
                        rf[me] = __fadd_rn(c, ew);
                        if (test_last && j == nb - 1 && fabs((double)ew) > (double)crtest)
                            bad = 1;
                    }
                }
                base += nchunk;
            }
            // border cells of every sweep but the block's last, each right after the interior cell it reads
            if (t < 6 * (nb - 1)) {
                const int j = t / 6, kind = t - 6 * j;
                const int d = tau - 2 * j;
                if (kind == 0) { // right column: (nx-1, y) from (nx-2, y)
                    const int y = d - (nx - 1);
                    if (y >= 1 && y <= ny - 2)
                        border(nx - 1, y, nx - 2, y);
                } else if (kind == 1) { // bottom row: (x, ny-1) from (x, ny-2), x >= 1 (the right corner reads the right column)
                    const int x = d - (ny - 1);
                    if (x >= 1 && x <= nx - 1)
                        border(x, ny - 1, x, ny - 2);
                } else if (kind == 2) { // left column: (0, y) from (1, y), two steps behind its diagonal
                    const int y = d - 2;
                    if (y >= 1 && y <= ny - 2)
                        border(0, y, 1, y);
                } else if (kind == 3) { // top row: (x, 0) from (x, 1), x >= 1
                    const int x = d - 2;
                    if (x >= 1 && x <= nx - 1)
                        border(x, 0, x, 1);
                } else if (kind == 4) { // corner (0, 0) from the left column's (0, 1)
                    if (d == 4)
                        border(0, 0, 0, 1);
                } else { // corner (0, ny-1) from the left column's (0, ny-2)
                    if (d == ny + 1)
                        border(0, ny - 1, 0, ny - 2);
                }
            }
            store_diag(tau - 2 * (nb - 1) - 5);
            __syncthreads();
        }
        for (int d = tau_end - 2 * (nb - 1) - 4; d <= dmax; ++d)
            store_diag(d);
        if (bad)
            s_bad = 1;
        __syncthreads();
        loop += nb;
        if (test_last && s_bad == 0)
            return; // convergence (:1334-1352)
        // the border pass of the block's last sweep, on global memory (:1354-1363)
        const int nxm1 = nx - 1, nym1 = ny - 1;
        for (int y = 1 + t; y < nym1; y += kT) {
            const size_t r = (size_t)y * nx;
            f[r] = __fadd_rn(f[r], __fmul_rn(__fsub_rn(f[r + 1], f[r]), w[r]));
            f[r + nxm1] = __fadd_rn(f[r + nxm1], __fmul_rn(__fsub_rn(f[r + nxm1 - 1], f[r + nxm1]), w[r + nxm1]));
        }
        __syncthreads();
        for (int x = t; x < nx; x += kT) {
            const size_t top = (size_t)x, bot = (size_t)nym1 * nx + x;
            f[top] = __fadd_rn(f[top], __fmul_rn(__fsub_rn(f[top + nx], f[top]), w[top]));
            f[bot] = __fadd_rn(f[bot], __fmul_rn(__fsub_rn(f[bot - nx], f[bot]), w[bot]));
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------- creepfill2d
// use_mean: mifi_creepfill2d_f (first guess = mean of the defined values, :1486-1507); else mifi_creepfillval2d_f (:1509-1525)
__global__ void __launch_bounds__(kT) k_creepfill2d(float* __restrict__ field, signed char* __restrict__ wfield,
                                                    unsigned short* __restrict__ rfield, int nx, int ny, int use_mean, float defaultVal,
                                                    unsigned short repeat, signed char setWeight, unsigned long long* __restrict__ n_changed)
{
    __shared__ __align__(16) float s_buf[kChunk];
    __shared__ unsigned long long s_changed;
    const size_t n = (size_t)nx * ny;
    float* f = field + blockIdx.x * n;
    signed char* w = wfield + blockIdx.x * n;
    unsigned short* r = rfield + blockIdx.x * n;
    const int t = threadIdx.x;
    const Stats st = level_sum(f, n, s_buf);
    if (t == 0 && n_changed)
        n_changed[blockIdx.x] = st.n_nan;
    const unsigned long long nUnchanged = n - st.n_nan;
    if (nUnchanged == 0 || st.n_nan == 0)
        return; // :1383-1387 (and :1505 for the mean variant)
    const float guess = use_mean ? (float)__ddiv_rn(st.sum, (double)nUnchanged) : defaultVal;
    for (size_t i = t; i < n; i += kT) { // :1409-1422
        if (isnan(f[i])) {
            w[i] = 0;
            r[i] = 0;
            f[i] = guess;
        } else {
            w[i] = setWeight;
            r[i] = repeat;
        }
    }
    if (t == 0)
        s_changed = 1;
    __syncthreads();
    const int nxm1 = nx - 1, nym1 = ny - 1;
    unsigned long long l = 0;
    while (true) { // :1431-1464
        const unsigned long long before = s_changed;
        __syncthreads();
        if (!(before > 0 && l < nUnchanged))
            break;
        if (t == 0)
            s_changed = 0;
        ++l;
        unsigned mine = 0;
        for (int d = 2; d <= nxm1 - 1 + nym1 - 1; ++d) {
            const int ylo = d - (nxm1 - 1) > 1 ? d - (nxm1 - 1) : 1;
            const int yhi = d - 1 < nym1 - 1 ? d - 1 : nym1 - 1;
            for (int y = ylo + t; y <= yhi; y += kT) {
                const size_t p = (size_t)y * nx + (d - y);
                if (r[p] < repeat) {
                    const int w1 = w[p + 1], w2 = w[p - 1], w3 = w[p + nx], w4 = w[p - nx];
                    const size_t wsum = (size_t)(long long)(w1 + w2 + w3 + w4); // chars promote to int, the sum converts to size_t
                    if (wsum != 0) {
                        const float acc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn((float)w1, f[p + 1]), __fmul_rn((float)w2, f[p - 1])),
                                                              __fmul_rn((float)w3, f[p + nx])),
                                                    __fmul_rn((float)w4, f[p - nx]));
                        const float v = __fadd_rn(f[p], acc);
                        f[p] = __fdiv_rn(v, (float)(1 + wsum));
                        w[p] = 1;
                        r[p] = (unsigned short)(r[p] + 1);
                        ++mine;
                    }
                }
            }
            __syncthreads();
        }
        if (mine)
            atomicAdd(&s_changed, (unsigned long long)mine);
        __syncthreads();
    }
    // simple calculations at the borders (:1466-1490)
    for (unsigned short k = 0; k < repeat; ++k) {
        for (int y = 1 + t; y < nym1; y += kT) {
            const size_t row = (size_t)y * nx;
            if (r[row] < repeat) {
                const float v = __fadd_rn(f[row], __fmul_rn(f[row + 1], (float)w[row + 1]));
                f[row] = __fdiv_rn(v, (float)(1 + w[row + 1]));
                w[row] = 1;
            }
            if (r[row + nxm1] < repeat) {
                const float v = __fadd_rn(f[row + nxm1], __fmul_rn(f[row + nxm1 - 1], (float)w[row + nxm1 - 1]));
                f[row + nxm1] = __fdiv_rn(v, (float)(1 + w[row + nxm1 - 1]));
                w[row + nxm1] = 1;
            }
        }
        __syncthreads();
        for (int x = t; x < nx; x += kT) {
            const size_t top = (size_t)x, bot = (size_t)nym1 * nx + x;
            if (r[top] < repeat) {
                const float v = __fadd_rn(f[top], __fmul_rn(f[top + nx], (float)w[top + nx]));
                f[top] = __fdiv_rn(v, (float)(1 + w[top + nx]));
                w[top] = 1;
            }
            if (r[bot] < repeat) {
                const float v = __fadd_rn(f[bot], __fmul_rn(f[bot - nx], (float)w[bot - nx]));
                f[bot] = __fdiv_rn(v, (float)(1 + w[bot - nx]));
                w[bot] = 1;
            }
        }
        __syncthreads();
    }
}

} // namespace

// nz levels of nx x ny in place on the device; d_nchanged (nz counters, may be null) receives the NaN count of every level
int launch_fill2d(float* d_field, size_t nx, size_t ny, size_t nz, float relaxCrit, float corrEff, size_t maxLoop,
                  unsigned long long* d_nchanged, cudaStream_t st)
{
    if (nx * ny == 0 || nz == 0)
        return FB_OK;
    FB_REQUIRE(nx >= 2 && ny >= 2 && nx < 2147483647u && ny < 2147483647u, "fill2d needs at least 2 x 2 points per level");
    float* d_w = nullptr;
    FB_CUDA_CHECK(cudaMallocAsync(&d_w, sizeof(float) * nx * ny * nz, st));
    // skewed (pipelined) sweeps when the diagonals in flight fit in shared memory: 32 diagonals (10 sweeps) or 16 (5 sweeps)
    const int by_y = ny <= nx ? 1 : 0;
    const int lp = (int)(by_y ? ny : nx) + 1;
    const size_t smem32 = sizeof(float) * 2 * 32 * (size_t)lp, smem16 = smem32 / 2;
    const size_t smem_max = 200 * 1024;
    cudaError_t ea = cudaSuccess;
    if (nx >= 3 && ny >= 3 && smem32 <= smem_max && !std::getenv("FIMEX_B200_FILL_SIMPLE")) {
        ea = cudaFuncSetAttribute(k_fill2d_skewed<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32);
        if (ea == cudaSuccess)
            k_fill2d_skewed<32><<<(unsigned)nz, kT, smem32, st>>>(d_field, d_w, (int)nx, (int)ny, by_y, lp, relaxCrit, corrEff,
                                                                 (unsigned long long)maxLoop, d_nchanged);
    } else if (nx >= 3 && ny >= 3 && smem16 <= smem_max && !std::getenv("FIMEX_B200_FILL_SIMPLE")) {
        ea = cudaFuncSetAttribute(k_fill2d_skewed<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16);
        if (ea == cudaSuccess)
            k_fill2d_skewed<16><<<(unsigned)nz, kT, smem16, st>>>(d_field, d_w, (int)nx, (int)ny, by_y, lp, relaxCrit, corrEff,
                                                                 (unsigned long long)maxLoop, d_nchanged);
    } else {
        k_fill2d<<<(unsigned)nz, kT, 0, st>>>(d_field, d_w, (int)nx, (int)ny, relaxCrit, corrEff, (unsigned long long)maxLoop, d_nchanged);
    }
    if (ea != cudaSuccess) {
        cudaFreeAsync(d_w, st);
        FB_CUDA_CHECK(ea);
    }
    count_launch();
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(d_w, st);
    FB_CUDA_CHECK(e);
    return FB_OK;
}

int launch_creepfill2d(float* d_field, size_t nx, size_t ny, size_t nz, bool use_mean, float defaultVal, unsigned short repeat,
                       signed char setWeight, unsigned long long* d_nchanged, cudaStream_t st)
{
    if (nx * ny == 0 || nz == 0)
        return FB_OK;
    FB_REQUIRE(nx >= 2 && ny >= 2 && nx < 2147483647u && ny < 2147483647u, "creepfill2d needs at least 2 x 2 points per level");
    signed char* d_w = nullptr;
    unsigned short* d_r = nullptr;
    FB_CUDA_CHECK(cudaMallocAsync(&d_w, nx * ny * nz, st));
    FB_CUDA_CHECK(cudaMallocAsync(&d_r, sizeof(unsigned short) * nx * ny * nz, st));
    k_creepfill2d<<<(unsigned)nz, kT, 0, st>>>(d_field, d_w, d_r, (int)nx, (int)ny, use_mean ? 1 : 0, defaultVal, repeat, setWeight, d_nchanged);
    count_launch();
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(d_w, st);
    cudaFreeAsync(d_r, st);
    FB_CUDA_CHECK(e);
    return FB_OK;
}

namespace {
// the reference's single-level host-pointer forms
int run_host_level(float* field, size_t nx, size_t ny, size_t* nChanged, int kind, float a, float b, size_t maxLoop, unsigned short repeat,
                   signed char setWeight)
{
    const size_t n = nx * ny;
    if (n == 0)
        return FB_OK; // :1248, :1380: *nChanged is left untouched
    cudaStream_t st = cudaStreamPerThread;
    float* d = nullptr;
    unsigned long long* d_n = nullptr;
    FB_CUDA_CHECK(cudaMallocAsync(&d, sizeof(float) * n, st));
    FB_CUDA_CHECK(cudaMallocAsync(&d_n, sizeof(unsigned long long), st));
    FB_CUDA_CHECK(cudaMemcpyAsync(d, field, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    int rc;
    if (kind == 0)
        rc = launch_fill2d(d, nx, ny, 1, a, b, maxLoop, d_n, st);
    else
        rc = launch_creepfill2d(d, nx, ny, 1, kind == 1, a, repeat, setWeight, d_n, st);
    unsigned long long h_n = 0;
    cudaError_t e = cudaSuccess;
    if (rc == FB_OK) {
        e = cudaMemcpyAsync(field, d, sizeof(float) * n, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(&h_n, d_n, sizeof(h_n), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(st);
    }
    cudaFreeAsync(d, st);
    cudaFreeAsync(d_n, st);
    if (rc != FB_OK)
        return rc;
    FB_CUDA_CHECK(e);
    if (nChanged)
        *nChanged = (size_t)h_n;
    return FB_OK;
}
} // namespace
} // namespace fb

extern "C" {

int mifi_fill2d_f(size_t nx, size_t ny, float* field, float relaxCrit, float corrEff, size_t maxLoop, size_t* nChanged)
{
    return fb::run_host_level(field, nx, ny, nChanged, 0, relaxCrit, corrEff, maxLoop, 0, 0);
}

int mifi_creepfill2d_f(size_t nx, size_t ny, float* field, unsigned short repeat, char setWeight, size_t* nChanged)
{
    return fb::run_host_level(field, nx, ny, nChanged, 1, 0.f, 0.f, 0, repeat, (signed char)setWeight);
}

int mifi_creepfillval2d_f(size_t nx, size_t ny, float* field, float defaultVal, unsigned short repeat, char setWeight, size_t* nChanged)
{
    return fb::run_host_level(field, nx, ny, nChanged, 2, defaultVal, 0.f, 0, repeat, (signed char)setWeight);
}

int fb200_fill2d_device(float* d_field, size_t nx, size_t ny, size_t nz, float relaxCrit, float corrEff, size_t maxLoop, void* cuda_stream)
{
    return fb::launch_fill2d(d_field, nx, ny, nz, relaxCrit, corrEff, maxLoop, nullptr,
                             cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : cudaStreamPerThread);
}

int fb200_creepfill2d_device(float* d_field, size_t nx, size_t ny, size_t nz, int useDefaultVal, float defaultVal, unsigned short repeat,
                             char setWeight, void* cuda_stream)
{
    return fb::launch_creepfill2d(d_field, nx, ny, nz, useDefaultVal == 0, defaultVal, repeat, (signed char)setWeight, nullptr,
                                  cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : cudaStreamPerThread);
}

} // extern "C"
