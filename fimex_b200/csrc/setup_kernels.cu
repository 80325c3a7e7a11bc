// fimex_b200/csrc/setup_kernels.cu -- once-per-grid kernels: K1 projection, K2 points2position, crop
// helpers, gather-table compilation, K7 rotation matrix.  fp64 throughout; not part of the throughput
// metric (amortised over every level of every slice) but kept on the device so the index
// tables never leave HBM.
//
// Where the result decides an integer (array index, table mode) the arithmetic is written with
// __dadd_rn/__dmul_rn/__ddiv_rn so that nvcc cannot contract it into FMAs: given identical inputs the
// positions and tables are bit-identical to what the reference's C code computes on x86-64.
#include "kernels.h"
#include "tables.cuh"

#include <math.h>

#include <vector>

namespace fb {

namespace {
constexpr int kThreads = 256;

inline int grid_for(long long n, int threads = kThreads)
{
    long long b = (n + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 32;
    if (b > cap)
        b = cap; // grid-stride loops below
    if (b < 1)
        b = 1;
    return (int)b;
}

// --------------------------------------------------------------------------------------------- K1
__global__ void k_project_mesh(ProjDef src, ProjDef dst, bool shift, const double* __restrict__ xaxis, const double* __restrict__ yaxis,
                               int nx, long long n, double* __restrict__ xo, double* __restrict__ yo, int* status)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double x = xaxis[i % nx];
        double y = yaxis[i / nx];
        const int e = transform_point(src, dst, shift, n == 1, x, y);
        if (e != 0)
            atomicCAS(status, 0, e);
        xo[i] = x;
        yo[i] = y;
    }
}

__global__ void k_project_values(ProjDef src, ProjDef dst, bool shift, long long n, double* __restrict__ xs, double* __restrict__ ys,
                                 int* status)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double x = xs[i];
        double y = ys[i];
        const int e = transform_point(src, dst, shift, n == 1, x, y);
        if (e != 0)
            atomicCAS(status, 0, e);
        xs[i] = x;
        ys[i] = y;
    }
}

// --------------------------------------------------------------------------------------------- K2
struct AxisInfo {
    int num;
    int ascending;
    int is_longitude;
    int fold_high; // axis spans negative longitudes: points > pi get -2pi; else negative points get +2pi
    int circular;
};

__global__ void k_points2position(double* __restrict__ pts, long long n, const double* __restrict__ axis, AxisInfo ai)
{
    const double two_pi = 2 * FB_PI;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double p = pts[i];
        if (ai.is_longitude) {
            if (ai.fold_high) {
                if (p > FB_PI)
                    p = __dsub_rn(p, two_pi);
            } else {
                if (p < 0)
                    p = __dadd_rn(p, two_pi);
            }
        }
        if (!isfinite(p)) {
            pts[i] = -999.;
            continue;
        }
        // the reference's bisection, probe for probe (bsearchDoubleIndex, interpolation.c:124-146)
        int first = 0, last = ai.num - 1, pos = 0, cmp = 0;
        while (first <= last) {
            pos = (first + last) / 2;
            const double b = axis[pos];
            cmp = (p > b) ? 1 : ((p == b) ? 0 : -1);
            if (!ai.ascending)
                cmp = -cmp;
            if (cmp > 0)
                first = pos + 1;
            else if (cmp < 0)
                last = pos - 1;
            else
                break;
        }
        if (cmp == 0) {
            pts[i] = (double)pos;
            continue;
        }
        int seg = (cmp > 0) ? pos + 1 : pos;
        if (seg == ai.num)
            seg--;
        else if (seg == 0)
            seg++;
        const double hi = axis[seg];
        const double slope = __dsub_rn(hi, axis[seg - 1]);
        const double offset = __dsub_rn(hi, __dmul_rn(slope, (double)seg));
        double apos = __ddiv_rn(__dsub_rn(p, offset), slope);
        if (ai.circular && apos <= -0.5)
            apos = __dadd_rn(apos, (double)ai.num);
        if (ai.circular && apos > __dsub_rn((double)ai.num, 0.5))
            apos = __dsub_rn(apos, (double)ai.num);
        pts[i] = apos;
    }
}

// --------------------------------------------------------------------------------------------- crop helpers
__global__ void k_minmax(const double* __restrict__ v, long long n, double* __restrict__ partial)
{
    __shared__ double smin[kThreads / 32], smax[kThreads / 32];
    double lo = v[0], hi = v[0];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = v[i];
        if (x < lo)
            lo = x;
        if (hi < x)
            hi = x;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double l2 = __shfl_xor_sync(0xffffffffu, lo, o);
        const double h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        if (l2 < lo)
            lo = l2;
        if (hi < h2)
            hi = h2;
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        smin[w] = lo;
        smax[w] = hi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < kThreads / 32; ++k) {
            if (smin[k] < lo)
                lo = smin[k];
            if (hi < smax[k])
                hi = smax[k];
        }
        partial[2 * blockIdx.x] = lo;
        partial[2 * blockIdx.x + 1] = hi;
    }
}

__global__ void k_shift(double* __restrict__ v, long long n, double delta)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        v[i] = __dsub_rn(v[i], delta);
}

__global__ void k_scale(double* __restrict__ v, long long n, double f)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        v[i] = __dmul_rn(f, v[i]);
}

// --------------------------------------------------------------------------------------------- table compilation
__global__ void k_compile_nn(const double* __restrict__ px, const double* __restrict__ py, long long n, int ix, int iy,
                             int* __restrict__ off)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = px[i], y = py[i];
        int o = -1;
        if (fits_int(x) && fits_int(y)) {
            const long long rx = llround(x), ry = llround(y); // half away from zero, like lround (:864-865)
            if (rx >= 0 && rx < ix && ry >= 0 && ry < iy)
                o = (int)(ry * ix + rx);
        }
        off[i] = o;
    }
}

__global__ void k_compile_bilinear(const double* __restrict__ px, const double* __restrict__ py, long long n, int ix, int iy,
                                   int4* __restrict__ tab)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        tab[i] = classify_bilinear(px[i], py[i], ix, iy);
}

__global__ void k_compile_bicubic(const double* __restrict__ px, const double* __restrict__ py, long long n, int ix, int iy,
                                  int* __restrict__ off, double2* __restrict__ frac)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = px[i], y = py[i];
        int o = -1;
        double2 f = make_double2(0., 0.);
        if (fits_int(x) && fits_int(y)) {
            const int x0 = (int)floor(x), y0 = (int)floor(y);
            if (1 <= x0 && x0 + 2 < ix && 1 <= y0 && y0 + 2 < iy) { // :975-976
                o = (y0 - 1) * ix + (x0 - 1);
                f.x = __dsub_rn(x, (double)x0);
                f.y = __dsub_rn(y, (double)y0);
            }
        }
        off[i] = o;
        frac[i] = f;
    }
}

__device__ __forceinline__ int round_and_clamp(double d, int maxi)
{
    // RoundAndClamp(0, maxi, -1): round() half away from zero, then range check (Utils.cc:42-58)
    const double r = round(d);
    if (!(r >= 0. && r <= (double)maxi))
        return -1;
    return (int)r;
}

__global__ void k_compile_forward_cells(const double* __restrict__ px, const double* __restrict__ py, long long n, int ox, int oy,
                                        int* __restrict__ cell)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int xi = round_and_clamp(px[i], ox - 1);
        const int yi = round_and_clamp(py[i], oy - 1);
        cell[i] = (xi >= 0 && yi >= 0) ? yi * ox + xi : -1;
    }
}

// --------------------------------------------------------------------------------------------- K7
__device__ __forceinline__ double bearing(double lat0, double lon0, double lat1, double lon1)
{
    // mifi_bearing, interpolation.c:311-329
    const double dlon = lon0 - lon1;
    double sd, cd, s0, c0, s1, c1;
    sincos(dlon, &sd, &cd);
    sincos(lat0, &s0, &c0);
    sincos(lat1, &s1, &c1);
    return atan2(sd * c1, c0 * s1 - s0 * c1 * cd);
}

__global__ void k_vector_matrix(ProjDef in, ProjDef out, bool shift, const double* __restrict__ in_x, const double* __restrict__ in_y,
                                const double* __restrict__ out_x, const double* __restrict__ out_y, double dx, double dy, long long n,
                                double* __restrict__ matrix, int* status)
{
    const bool ll = out.is_latlong != 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x0 = in_x[i], y0 = in_y[i];
        const double ox = out_x[i], oy = out_y[i];
        // step along x (:350-381)
        double tx = x0 + dx, ty = y0;
        int e = transform_point(in, out, shift, n == 1, tx, ty);
        if (e != 0)
            atomicCAS(status, 0, e);
        double phiy;
        if (ll) {
            phiy = bearing(oy, ox, ty, tx);
        } else {
            phiy = atan2(ty - oy, tx - ox);
            if (!(dx > 0))
                phiy += FB_PI;
        }
        // step along y (:391-433)
        tx = x0;
        ty = y0 + dy;
        e = transform_point(in, out, shift, n == 1, tx, ty);
        if (e != 0)
            atomicCAS(status, 0, e);
        double phi0;
        if (ll) {
            phi0 = bearing(oy, ox, ty, tx);
            if (!(dy > 0))
                phi0 += FB_PI;
        } else {
            double phix = -1 * atan2(tx - ox, ty - oy);
            if (!(dy > 0))
                phix += FB_PI;
            phi0 = .5 * (phix + phiy); // no wrap handling, as :424
        }
        double s, c;
        sincos(phi0, &s, &c);
        matrix[4 * i + 0] = c;
        matrix[4 * i + 1] = s;
        matrix[4 * i + 2] = -1 * s;
        matrix[4 * i + 3] = phi0;
    }
}

__global__ void k_matrix_to_cossin(const double* __restrict__ m, long long n, double2* __restrict__ cs)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        cs[i] = make_double2(m[4 * i], m[4 * i + 1]);
}

} // namespace

// ------------------------------------------------------------------------------------------------- launchers
int launch_project_mesh(const ProjDef& src, const ProjDef& dst, const double* d_xaxis, const double* d_yaxis, int nx, int ny, double* d_xo,
                        double* d_yo, int* d_status, cudaStream_t st)
{
    const long long n = (long long)nx * ny;
    if (n == 0)
        return FB_OK;
    k_project_mesh<<<grid_for(n), kThreads, 0, st>>>(src, dst, needs_datum_shift(src, dst), d_xaxis, d_yaxis, nx, n, d_xo, d_yo, d_status);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_project_values(const ProjDef& src, const ProjDef& dst, double* d_x, double* d_y, long long n, int* d_status, cudaStream_t st)
{
    if (n == 0)
        return FB_OK;
    k_project_values<<<grid_for(n), kThreads, 0, st>>>(src, dst, needs_datum_shift(src, dst), n, d_x, d_y, d_status);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_points2position(double* d_points, long long n, const double* d_axis, const double* h_axis, int num, int axis_type,
                           cudaStream_t st)
{
    if (n == 0)
        return FB_OK;
    FB_REQUIRE(num >= 2, "mifi_points2position: axis needs at least 2 values");
    AxisInfo ai;
    ai.num = num;
    ai.ascending = h_axis[0] < h_axis[num - 1];
    ai.is_longitude = (axis_type == FB_AXIS_LONGITUDE);
    ai.fold_high = 0;
    ai.circular = 0;
    if (ai.is_longitude) { // interpolation.c:155-180, evaluated on the host in the same fp64 arithmetic
        ai.fold_high = (h_axis[0] < 0 || h_axis[num - 1] < 0);
        volatile double next = h_axis[num - 1] + (h_axis[1] - h_axis[0]) * 1.01;
        if (ai.ascending) {
            next -= 2 * FB_PI;
            ai.circular = (next >= h_axis[0]);
        } else {
            next += 2 * FB_PI;
            ai.circular = (next <= h_axis[0]);
        }
    }
    k_points2position<<<grid_for(n), kThreads, 0, st>>>(d_points, n, d_axis, ai);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int device_minmax(const double* d_v, long long n, double* h_min, double* h_max, cudaStream_t st)
{
    FB_REQUIRE(n > 0, "minmax of an empty array");
    int blocks = grid_for(n);
    if (blocks > 1024)
        blocks = 1024;
    double* d_partial = nullptr;
    FB_CUDA_CHECK(cudaMallocAsync(&d_partial, sizeof(double) * 2 * blocks, st));
    k_minmax<<<blocks, kThreads, 0, st>>>(d_v, n, d_partial);
    count_launch();
    std::vector<double> h(2 * (size_t)blocks);
    FB_CUDA_CHECK(cudaMemcpyAsync(h.data(), d_partial, sizeof(double) * 2 * blocks, cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    FB_CUDA_CHECK(cudaFreeAsync(d_partial, st));
    double lo = h[0], hi = h[1];
    for (int b = 1; b < blocks; ++b) {
        if (h[2 * b] < lo)
            lo = h[2 * b];
        if (hi < h[2 * b + 1])
            hi = h[2 * b + 1];
    }
    *h_min = lo;
    *h_max = hi;
    return FB_OK;
}

int launch_shift(double* d_v, long long n, double delta, cudaStream_t st)
{
    if (n == 0)
        return FB_OK;
    k_shift<<<grid_for(n), kThreads, 0, st>>>(d_v, n, delta);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_scale(double* d_v, long long n, double factor, cudaStream_t st)
{
    if (n == 0)
        return FB_OK;
    k_scale<<<grid_for(n), kThreads, 0, st>>>(d_v, n, factor);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_compile_nn(const double* d_px, const double* d_py, long long n, int ix, int iy, int* d_off, cudaStream_t st)
{
    if (n == 0)
        return FB_OK;
    k_compile_nn<<<grid_for(n), kThreads, 0, st>>>(d_px, d_py, n, ix, iy, d_off);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_compile_bilinear(const double* d_px, const double* d_py, long long n, int ix, int iy, int4* d_tab, cudaStream_t st)
{
    if (n == 0)
        return FB_OK;
    k_compile_bilinear<<<grid_for(n), kThreads, 0, st>>>(d_px, d_py, n, ix, iy, d_tab);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_compile_bicubic(const double* d_px, const double* d_py, long long n, int ix, int iy, int* d_off, double2* d_frac,
                           cudaStream_t st)
{
    if (n == 0)
        return FB_OK;
    k_compile_bicubic<<<grid_for(n), kThreads, 0, st>>>(d_px, d_py, n, ix, iy, d_off, d_frac);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_compile_forward_cells(const double* d_px, const double* d_py, long long n, int ox, int oy, int* d_cell, cudaStream_t st)
{
    if (n == 0)
        return FB_OK;
    k_compile_forward_cells<<<grid_for(n), kThreads, 0, st>>>(d_px, d_py, n, ox, oy, d_cell);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_vector_matrix(const ProjDef& in, const ProjDef& out, const double* d_in_x, const double* d_in_y, const double* d_out_x,
                         const double* d_out_y, double dx, double dy, long long n, double* d_matrix, int* d_status, cudaStream_t st)
{
    if (n == 0)
        return FB_OK;
    k_vector_matrix<<<grid_for(n), kThreads, 0, st>>>(in, out, needs_datum_shift(in, out), d_in_x, d_in_y, d_out_x, d_out_y, dx, dy, n,
                                                      d_matrix, d_status);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_matrix_to_cossin(const double* d_matrix, long long n, double2* d_cs, cudaStream_t st)
{
    if (n == 0)
        return FB_OK;
    k_matrix_to_cossin<<<grid_for(n), kThreads, 0, st>>>(d_matrix, n, d_cs);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

} // namespace fb
