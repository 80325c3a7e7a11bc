// Cost of a 128-bit shared load when only part of the warp's lanes are active (predicated off per lane), and when lanes share
// addresses.  16 warps per SM, 8 independent loads per iteration into separate integer accumulators (no dependent FP chain).
// Addresses: lane L reads the 16 bytes at word offset slot(L) * 12 (the tap-major staging stride of the gathers: consecutive
// slots never share a bank group).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_pred lds_pred.cu && ./lds_pred
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// SLOTDIV: lanes L and L' read the same address when L / SLOTDIV == L' / SLOTDIV
__global__ void k(unsigned* out, int iters, long long* cyc, unsigned mask, int slotdiv)
{
    __shared__ __align__(16) unsigned s[12288];
    for (int i = threadIdx.x; i < 12288; i += blockDim.x) s[i] = i * 2654435761u;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool on = (mask >> lane) & 1u;
    const uint4* p = reinterpret_cast<const uint4*>(s) + (lane / slotdiv) * 3 + w; // 3 uint4 = 12 words per slot
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0, a6 = a0, a7 = a0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (on) {
            const uint4 v0 = p[0], v1 = p[128], v2 = p[256], v3 = p[384], v4 = p[512], v5 = p[640], v6 = p[768], v7 = p[896];
            a0.x ^= v0.x, a0.y ^= v0.y, a0.z ^= v0.z, a0.w ^= v0.w;
            a1.x ^= v1.x, a1.y ^= v1.y, a1.z ^= v1.z, a1.w ^= v1.w;
            a2.x ^= v2.x, a2.y ^= v2.y, a2.z ^= v2.z, a2.w ^= v2.w;
            a3.x ^= v3.x, a3.y ^= v3.y, a3.z ^= v3.z, a3.w ^= v3.w;
            a4.x ^= v4.x, a4.y ^= v4.y, a4.z ^= v4.z, a4.w ^= v4.w;
            a5.x ^= v5.x, a5.y ^= v5.y, a5.z ^= v5.z, a5.w ^= v5.w;
            a6.x ^= v6.x, a6.y ^= v6.y, a6.z ^= v6.z, a6.w ^= v6.w;
            a7.x ^= v7.x, a7.y ^= v7.y, a7.z ^= v7.z, a7.w ^= v7.w;
        }
        p += (a0.x == 0x12345u); // keeps the loads inside the loop
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    uint4 r = a0;
#define FOLD(A) r.x ^= A.x * 3u, r.y ^= A.y * 5u, r.z ^= A.z * 7u, r.w ^= A.w * 11u;
    FOLD(a1) FOLD(a2) FOLD(a3) FOLD(a4) FOLD(a5) FOLD(a6) FOLD(a7)
    out[blockIdx.x * blockDim.x + threadIdx.x] = r.x + 13u * r.y + 17u * r.z + 19u * r.w;
}
int run(const char* name, unsigned mask, int slotdiv = 1)
{
    unsigned* out; long long* cyc; CK(cudaMalloc(&out, 148 * 1024 * 4)); CK(cudaMalloc(&cyc, 148 * 8));
    const int iters = 4000, threads = 512;
    k<<<148, threads>>>(out, iters, cyc, mask, slotdiv);
    k<<<148, threads>>>(out, iters, cyc, mask, slotdiv);
    CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("%-58s mask %08x  %6.2f cycles per warp load instruction\n", name, mask, (double)h[0] / (iters * 8.0 * (threads / 32)));
    cudaFree(out); cudaFree(cyc);
    return 0;
}
int main()
{
    printf("LDS.128, 16 warps per SM, 8 independent loads per iteration; distinct slots never share a bank group\n");
    run("all 32 lanes, 32 distinct addresses", 0xffffffffu);
    run("16 lanes (every other)", 0x55555555u);
    run("16 lanes (lower half)", 0x0000ffffu);
    run("8 lanes (every 4th)", 0x11111111u);
    run("8 lanes (one quarter warp)", 0x000000ffu);
    run("8 lanes (irregular)", 0x40921084u);
    run("4 lanes (every 8th)", 0x01010101u);
    run("2 lanes", 0x00010001u);
    run("1 lane", 0x00000100u);
    run("no lane (whole warp predicated off / branched over)", 0x0u);
    run("all lanes, pairs of neighbours share (16 distinct)", 0xffffffffu, 2);
    run("all lanes, quads of neighbours share (8 distinct)", 0xffffffffu, 4);
    run("all lanes, groups of 3 share (11 distinct, irregular)", 0xffffffffu, 3);
    run("all lanes, groups of 5 share (7 distinct, irregular)", 0xffffffffu, 5);
    run("all lanes, one address", 0xffffffffu, 32);
    return 0;
}
