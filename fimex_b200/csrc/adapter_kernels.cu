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
