// store-bandwidth ceilings on B200 for the access patterns of the gather kernels (no loads, no math)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_linear(float4* out, size_t n4)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        __stcs(out + i, make_float4(1.f, 2.f, 3.f, 4.f));
}

// CTA = tile of TX x TY points of a level of ox x oy, 256 threads, lane = x (TX/32 segments per row), chunk of levels
template <int TX, int TY, int VEC>
__global__ void k_tiles(float* out, int ox, int oy, int nz, int chunk)
{
    const int tiles_x = ox / TX;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int z0 = blockIdx.y * chunk, z1 = min(nz, z0 + chunk);
    constexpr int per_row = TX / VEC;            // threads per tile row
    constexpr int rows_per_pass = 256 / per_row; // rows written per pass of the CTA
    const int lx = (threadIdx.x % per_row) * VEC, ly = threadIdx.x / per_row;
    const size_t level = (size_t)ox * oy;
    for (int z = z0; z < z1; ++z) {
        float* base = out + z * level + (size_t)(ty * TY) * ox + tx * TX + lx;
#pragma unroll
        for (int r = ly; r < TY; r += rows_per_pass) {
            if (VEC == 4)
                __stcs(reinterpret_cast<float4*>(base + (size_t)r * ox), make_float4(1.f, 2.f, 3.f, (float)z));
            else if (VEC == 2)
                __stcs(reinterpret_cast<float2*>(base + (size_t)r * ox), make_float2(1.f, (float)z));
            else
                __stcs(base + (size_t)r * ox, (float)z);
        }
    }
}


// bilinear with an in-register 4x4 transpose: warp = 32 columns x 4 rows (rows 4 apart), lane L writes a float4 at columns
// 4*(L>>2).., row (L&3): one STG.128 covers 4 rows x 128 B
template <int TX, int TY>
__global__ void k_tiles_tr(float* out, int ox, int oy, int nz, int chunk)
{
    const int tiles_x = ox / TX;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int z0 = blockIdx.y * chunk, z1 = min(nz, z0 + chunk);
    const int w = threadIdx.x >> 5, L = threadIdx.x & 31;
    constexpr int segs = TX / 32;                 // warps side by side
    const int lx = (w % segs) * 32 + 4 * (L >> 2);
    const int ly = (w / segs) + (8 / segs) * (L & 3) ; // rows RowStep apart, RowStep = 8/segs
    const size_t level = (size_t)ox * oy;
    for (int z = z0; z < z1; ++z) {
        float* base = out + z * level + (size_t)(ty * TY) * ox + tx * TX + lx;
#pragma unroll
        for (int r = ly; r < TY; r += 4 * (8 / segs))
            __stcs(reinterpret_cast<float4*>(base + (size_t)r * ox), make_float4(1.f, 2.f, 3.f, (float)z));
    }
}


// like k_tiles<TX, TY, VEC>, but every row of the tile is shifted left so that the warp stores start on 128-byte boundaries
template <int TX, int TY, int VEC>
__global__ void k_tiles_al(float* out, int ox, int oy, int nz, int chunk)
{
    const int tiles_x = (ox + 31 + TX - 1) / TX;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int z0 = blockIdx.y * chunk, z1 = min(nz, z0 + chunk);
    constexpr int per_row = TX / VEC;
    constexpr int rows_per_pass = 256 / per_row;
    const int lx = (threadIdx.x % per_row) * VEC, ly = threadIdx.x / per_row;
    const size_t level = (size_t)ox * oy;
    for (int z = z0; z < z1; ++z) {
#pragma unroll
        for (int r = ly; r < TY; r += rows_per_pass) {
            const int y = ty * TY + r;
            const int x = tx * TX + lx - (int)(((size_t)y * ox) & 31);
            if (x >= 0 && x + VEC <= ox && y < oy) {
                float* p = out + z * level + (size_t)y * ox + x;
                if (VEC == 4)
                    __stcs(reinterpret_cast<float4*>(p), make_float4(1.f, 2.f, 3.f, (float)z));
                else
                    __stcs(p, (float)z);
            }
        }
    }
}


// k_tiles<TX, TY, 1> launched as thread-block clusters of CL x-adjacent tiles (co-scheduled: their row segments are written together)
template <int TX, int TY, int CL>
__global__ void __cluster_dims__(CL, 1, 1) k_tiles_cl(float* out, int ox, int oy, int nz, int chunk, int tiles_x)
{
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int z0 = blockIdx.y * chunk, z1 = min(nz, z0 + chunk);
    constexpr int rows_per_pass = 256 / TX;
    const int lx = threadIdx.x % TX, ly = threadIdx.x / TX;
    const size_t level = (size_t)ox * oy;
    for (int z = z0; z < z1; ++z) {
        float* base = out + z * level + (size_t)(ty * TY) * ox + tx * TX + lx;
#pragma unroll
        for (int r = ly; r < TY; r += rows_per_pass)
            __stcs(base + (size_t)r * ox, (float)z);
    }
}


// k_tiles<TX, TY, 1> with other cache operators on the store: 0 = default (write-back, evict-normal), 1 = .cg, 2 = .wt, 3 = .cs
template <int TX, int TY, int OP>
__global__ void k_tiles_op(float* out, int ox, int oy, int nz, int chunk)
{
    const int tiles_x = ox / TX;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int z0 = blockIdx.y * chunk, z1 = min(nz, z0 + chunk);
    constexpr int rows_per_pass = 256 / TX;
    const int lx = threadIdx.x % TX, ly = threadIdx.x / TX;
    const size_t level = (size_t)ox * oy;
    for (int z = z0; z < z1; ++z) {
        float* base = out + z * level + (size_t)(ty * TY) * ox + tx * TX + lx;
#pragma unroll
        for (int r = ly; r < TY; r += rows_per_pass) {
            float* p = base + (size_t)r * ox;
            if (OP == 0) *p = (float)z;
            else if (OP == 1) __stcg(p, (float)z);
            else if (OP == 2) __stwt(p, (float)z);
            else __stcs(p, (float)z);
        }
    }
}

template <class F>
float time_ms(F f, int reps = 5)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char** argv)
{
    const int ox = argc > 1 ? atoi(argv[1]) : 2048, oy = 2016, nz = 1644; // row pitch 4*ox bytes: 2000 -> rows start on 64-byte, not 128-byte, boundaries
    printf("ox = %d (row pitch %d B)\n", ox, 4 * ox); // 2048 x 2016 x 1644 floats = 27 GB
    const size_t n = (size_t)ox * oy * nz;
    float* out; CK(cudaMalloc(&out, n * 4));
    float* in; CK(cudaMalloc(&in, n * 4));
    const double gb = n * 4 / 1e9;
    float ms = time_ms([&] { k_linear<<<148 * 16, 256>>>((float4*)out, n / 4); });
    printf("linear float4 stores        %8.3f ms %8.1f GB/s\n", ms, gb / ms * 1e3);
    ms = time_ms([&] { CK(cudaMemsetAsync(out, 0, n * 4)); });
    printf("cudaMemset                  %8.3f ms %8.1f GB/s\n", ms, gb / ms * 1e3);
    ms = time_ms([&] { CK(cudaMemcpyAsync(out, in, n * 4, cudaMemcpyDeviceToDevice)); });
    printf("cudaMemcpy D2D (r+w bytes)  %8.3f ms %8.1f GB/s\n", ms, 2 * gb / ms * 1e3);
    for (int chunk : {64, 16, 256}) {
        { dim3 g((ox / 32) * (oy / 32), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles<32, 32, 1><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 32x32  scalar chunk %3d %8.3f ms %8.1f GB/s\n", chunk, ms, gb / ms * 1e3); }
        { dim3 g((ox / 64) * (oy / 16), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles<64, 16, 1><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 64x16  scalar chunk %3d %8.3f ms %8.1f GB/s\n", chunk, ms, gb / ms * 1e3); }
        { dim3 g((ox / 128) * (oy / 8), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles<128, 8, 1><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 128x8  scalar chunk %3d %8.3f ms %8.1f GB/s\n", chunk, ms, gb / ms * 1e3); }
        { dim3 g((ox / 128) * (oy / 8), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles<128, 8, 4><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 128x8  float4 chunk %3d %8.3f ms %8.1f GB/s\n", chunk, ms, gb / ms * 1e3); }
        { dim3 g((ox / 32) * (oy / 32), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles<32, 32, 4><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 32x32  float4 chunk %3d %8.3f ms %8.1f GB/s\n", chunk, ms, gb / ms * 1e3); }
        { dim3 g((ox / 64) * (oy / 16), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles<64, 16, 4><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 64x16  float4 chunk %3d %8.3f ms %8.1f GB/s\n", chunk, ms, gb / ms * 1e3); }
        { dim3 g((ox / 64) * (oy / 16), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles_tr<64, 16><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 64x16  transp chunk %3d %8.3f ms %8.1f GB/s\n", chunk, ms, gb / ms * 1e3); }
        { dim3 g((ox / 128) * (oy / 8), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles_tr<128, 8><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 128x8  transp chunk %3d %8.3f ms %8.1f GB/s\n", chunk, ms, gb / ms * 1e3); }
        { dim3 g((ox / 256) * (oy / 4), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles<256, 4, 1><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 256x4  scalar chunk %3d %8.3f ms %8.1f GB/s\n", chunk, ms, gb / ms * 1e3); }
        { dim3 g((ox / 256) * (oy / 4), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles<256, 4, 4><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 256x4  float4 chunk %3d %8.3f ms %8.1f GB/s\n", chunk, ms, gb / ms * 1e3); }
        { dim3 g((ox / 64) * (oy / 16), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles<64, 16, 2><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 64x16  float2 chunk %3d %8.3f ms %8.1f GB/s\n", chunk, ms, gb / ms * 1e3); }
        { dim3 g((ox / 128) * (oy / 8), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles<128, 8, 2><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 128x8  float2 chunk %3d %8.3f ms %8.1f GB/s\n", chunk, ms, gb / ms * 1e3); }
        { dim3 g(((ox + 31 + 63) / 64) * (oy / 16), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles_al<64, 16, 1><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 64x16  scalar ALIGNED %3d %8.3f ms %8.1f GB/s\n", chunk, ms, gb / ms * 1e3); }
        { dim3 g(((ox + 31 + 127) / 128) * (oy / 8), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles_al<128, 8, 4><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 128x8  float4 ALIGNED %3d %8.3f ms %8.1f GB/s\n", chunk, ms, gb / ms * 1e3); }
        { const int tcx = (ox / 64) / 4 * 4; dim3 g(tcx * (oy / 16), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles_cl<64, 16, 1><<<g, 256>>>(out, ox, oy, nz, chunk, tcx); });
          printf("tiles 64x16  scalar CLUSTER1 %3d %8.3f ms %8.1f GB/s (of %d columns)\n", chunk, ms, 4e-9 * tcx * 64 * (double)oy * nz / ms * 1e3, tcx * 64); }
        { const int tcx = (ox / 64) / 4 * 4; dim3 g(tcx * (oy / 16), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles_cl<64, 16, 4><<<g, 256>>>(out, ox, oy, nz, chunk, tcx); });
          printf("tiles 64x16  scalar CLUSTER4 %3d %8.3f ms %8.1f GB/s (of %d columns)\n", chunk, ms, 4e-9 * tcx * 64 * (double)oy * nz / ms * 1e3, tcx * 64); }
        { const int tcx = (ox / 64) / 2 * 2; dim3 g(tcx * (oy / 16), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles_cl<64, 16, 2><<<g, 256>>>(out, ox, oy, nz, chunk, tcx); });
          printf("tiles 64x16  scalar CLUSTER2 %3d %8.3f ms %8.1f GB/s (of %d columns)\n", chunk, ms, 4e-9 * tcx * 64 * (double)oy * nz / ms * 1e3, tcx * 64); }
        { dim3 g((ox / 64) * (oy / 16), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles_op<64, 16, 0><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 64x16  scalar st.wb     %3d %8.3f ms %8.1f GB/s\n", chunk, ms, 4e-9 * (ox / 64 * 64) * (double)oy * nz / ms * 1e3);
          ms = time_ms([&] { k_tiles_op<64, 16, 1><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 64x16  scalar st.cg     %3d %8.3f ms %8.1f GB/s\n", chunk, ms, 4e-9 * (ox / 64 * 64) * (double)oy * nz / ms * 1e3);
          ms = time_ms([&] { k_tiles_op<64, 16, 2><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 64x16  scalar st.wt     %3d %8.3f ms %8.1f GB/s\n", chunk, ms, 4e-9 * (ox / 64 * 64) * (double)oy * nz / ms * 1e3);
          ms = time_ms([&] { k_tiles_op<64, 16, 3><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("tiles 64x16  scalar st.cs     %3d %8.3f ms %8.1f GB/s\n", chunk, ms, 4e-9 * (ox / 64 * 64) * (double)oy * nz / ms * 1e3); }
    }
    return 0;
}
