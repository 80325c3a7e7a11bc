// shared-load throughput on B200 by access width and number of distinct addresses per warp
#include <cuda_runtime.h>
#include <cstdio>
template <int W, int DIV>   // W = words per lane (1, 2, 4); lanes L and L' read the same address when L/DIV == L'/DIV
__global__ void k(float* out, int iters, long long* cyc)
{
    __shared__ __align__(16) float s[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int stride = (W == 4) ? 12 : (W == 2 ? 10 : 9);     // tap-major strides that keep distinct taps in distinct banks
    int idx = ((lane / DIV) * stride + w * 4 * W) & 4095;
    idx = idx / W * W;
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int a = (idx + u * 512) & 8191;
            if (W == 4) { float4 v = *reinterpret_cast<const float4*>(s + a); acc += v.x + v.y + v.z + v.w; }
            else if (W == 2) { float2 v = *reinterpret_cast<const float2*>(s + a); acc += v.x + v.y; }
            else acc += s[a];
        }
        idx = (idx + (int)(acc == 12345.f)) & 4095; idx = idx / W * W;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int W, int DIV>
void run(const char* name)
{
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4000, threads = 512;
    k<W, DIV><<<148, threads>>>(out, iters, cyc);
    k<W, DIV><<<148, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = (double)h[0] / (iters * 8.0 * (threads / 32));
    printf("%-40s %6.2f cycles per warp load  = %6.1f B/clk/SM delivered to lanes (%d B per lane)\n", name, c, 32.0 * 4 * W / c, 4 * W);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<1, 1>("LDS.32  32 distinct");  run<1, 2>("LDS.32  16 distinct"); run<1, 32>("LDS.32  1 distinct");
    run<2, 1>("LDS.64  32 distinct");  run<2, 2>("LDS.64  16 distinct"); run<2, 4>("LDS.64  8 distinct"); run<2, 32>("LDS.64  1 distinct");
    run<4, 1>("LDS.128 32 distinct");  run<4, 2>("LDS.128 16 distinct"); run<4, 4>("LDS.128 8 distinct"); run<4, 8>("LDS.128 4 distinct"); run<4, 32>("LDS.128 1 distinct");
    return 0;
}
