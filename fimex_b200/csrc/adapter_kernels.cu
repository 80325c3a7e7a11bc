// fimex_b200/csrc/adapter_kernels.cu -- K10: the bad<->NaN adapters that sit on either side of the gather
// (CDMInterpolator::data2InterpolationArray / interpolationArray2Data, /root/reference/src/CDMInterpolator.cc:115-124;
// mifi_bad2nanf / mifi_nanf2bad, src/interpolation.c:1775-1793).
#include "../../include/fimex_b200.h"

#include "kernels.h"

namespace fb {
namespace {
constexpr int kThreads = 256;

__global__ void k_bad2nan(float* __restrict__ v, long long n, float bad)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = v[i];
        if (x == bad)
            v[i] = undef_f();
    }
}

__global__ void k_nan2bad(float* __restrict__ v, long long n, float bad)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = v[i];
        if (isnan(x))
            v[i] = bad;
    }
}

// data2InterpolationArray for a whole slab: asFloat() + mifi_bad2nanf in one pass (any CDM numeric type in, float out)
template <class T>
__global__ void k_as_float(const T* __restrict__ in, long long n, int has_bad, float bad, float* __restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = load_as_float<T>(in[i], has_bad != 0, bad);
}

// interpolationArray2Data for a whole slab (used where the gather kernel does not convert while storing).  No __restrict__:
// for float output the pass runs IN PLACE (in == out, run_slice_device); every thread reads element i before it writes it.
template <class Out>
__global__ void k_from_float(const float* in, long long n, Out conv, typename Out::type* out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = conv(in[i]);
}

int stream_blocks(long long n)
{
    long long blocks = (n + kThreads - 1) / kThreads;
    if (blocks > (long long)sm_count() * 32)
        blocks = (long long)sm_count() * 32;
    return (int)(blocks > 0 ? blocks : 1);
}

int run_inplace(float* pos, float* end, float bad, bool to_nan)
{
    if (isnan(bad) || pos == nullptr || end <= pos)
        return FB_OK; // a NaN badVal disables the pass (interpolation.c:1776, :1786)
    const long long n = end - pos;
    int dev = 0;
    FB_CUDA_CHECK(cudaGetDevice(&dev));
    cudaStream_t st = cudaStreamPerThread;
    float* d = nullptr;
    FB_CUDA_CHECK(cudaMallocAsync(&d, sizeof(float) * n, st));
    FB_CUDA_CHECK(cudaMemcpyAsync(d, pos, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    long long blocks = (n + kThreads - 1) / kThreads;
    if (blocks > (long long)sm_count() * 32)
        blocks = (long long)sm_count() * 32;
    if (to_nan)
        k_bad2nan<<<(int)blocks, kThreads, 0, st>>>(d, n, bad);
    else
        k_nan2bad<<<(int)blocks, kThreads, 0, st>>>(d, n, bad);
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    FB_CUDA_CHECK(cudaMemcpyAsync(pos, d, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaFreeAsync(d, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    return FB_OK;
}
} // namespace

int launch_as_float(int in_type, const void* d_in, long long n, bool has_bad, float bad, float* d_out, cudaStream_t st)
{
    if (n == 0)
        return FB_OK;
    const int blocks = stream_blocks(n);
    switch (in_type) {
#define FB_CASE(TAG, T)                                                                                                                    \
    case TAG:                                                                                                                              \
        k_as_float<T><<<blocks, kThreads, 0, st>>>(static_cast<const T*>(d_in), n, has_bad ? 1 : 0, bad, d_out);                           \
        break;
        FB_CASE(FB_T_CHAR, signed char)
        FB_CASE(FB_T_SHORT, short)
        FB_CASE(FB_T_INT, int)
        FB_CASE(FB_T_FLOAT, float)
        FB_CASE(FB_T_DOUBLE, double)
        FB_CASE(FB_T_UCHAR, unsigned char)
        FB_CASE(FB_T_USHORT, unsigned short)
        FB_CASE(FB_T_UINT, unsigned int)
        FB_CASE(FB_T_INT64, long long)
        FB_CASE(FB_T_UINT64, unsigned long long)
#undef FB_CASE
    default:
        FB_REQUIRE(false, "unsupported input data type " + std::to_string(in_type));
    }
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

int launch_from_float(const float* d_in, long long n, int out_type, double fill, void* d_out, cudaStream_t st)
{
    if (n == 0)
        return FB_OK;
    const int blocks = stream_blocks(n);
    switch (out_type) {
#define FB_CASE(TAG, T)                                                                                                                    \
    case TAG:                                                                                                                              \
        k_from_float<StoreAs<T>><<<blocks, kThreads, 0, st>>>(d_in, n, StoreAs<T>{cast_fill<T>(fill)}, static_cast<T*>(d_out));           \
        break;
        FB_CASE(FB_T_CHAR, signed char)
        FB_CASE(FB_T_SHORT, short)
        FB_CASE(FB_T_INT, int)
        FB_CASE(FB_T_FLOAT, float)
        FB_CASE(FB_T_DOUBLE, double)
        FB_CASE(FB_T_UCHAR, unsigned char)
        FB_CASE(FB_T_USHORT, unsigned short)
        FB_CASE(FB_T_UINT, unsigned int)
        FB_CASE(FB_T_INT64, long long)
        FB_CASE(FB_T_UINT64, unsigned long long)
#undef FB_CASE
    default:
        FB_REQUIRE(false, "unsupported output data type " + std::to_string(out_type));
    }
    count_launch();
    FB_CUDA_CHECK(cudaGetLastError());
    return FB_OK;
}

} // namespace fb

extern "C" {

size_t mifi_bad2nanf(float* posPtr, float* endPtr, float badVal)
{
    fb::run_inplace(posPtr, endPtr, badVal, true);
    return 0; // the reference always returns 0
}

size_t mifi_nanf2bad(float* posPtr, float* endPtr, float badVal)
{
    fb::run_inplace(posPtr, endPtr, badVal, false);
    return 0;
}

} // extern "C"
