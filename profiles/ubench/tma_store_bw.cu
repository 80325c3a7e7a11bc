// Store-bandwidth ceilings on B200 for output tiles staged in shared memory and written with TMA (cp.async.bulk.tensor
// shared -> global, 3-D box x * y * levels over the [z][y][x] output), next to the per-thread STG patterns of store_bw.cu.
// Also: the cost of a 128-bit shared load when only some lanes are active (the per-lane tap re-load experiment).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_store_bw tma_store_bw.cu && ./tma_store_bw [ox]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled encoder()
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("cuTensorMapEncodeTiled not found\n"); exit(1); }
    return (EncodeTiled)fn;
}

static CUtensorMap make_map(float* out, int ox, int oy, int nz, int bx, int by, int bz)
{
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)ox, (cuuint64_t)oy, (cuuint64_t)nz};
    cuuint64_t strides[2] = {(cuuint64_t)ox * 4, (cuuint64_t)ox * oy * 4};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
                 "r"((unsigned)__cvta_generic_to_shared(smem)), "r"(x), "r"(y), "r"(z)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gmem, const void* smem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem), "r"((unsigned)__cvta_generic_to_shared(smem)), "r"(bytes)
                 : "memory");
}

// CTA = tile TX x TY, 256 threads; per batch of LZ levels every thread writes its share of the TX*TY*LZ staged values to shared
// memory (lane = x, conflict-free), then one thread issues ONE tensor store of the whole box.  Two buffers.
// MODE 0: tensor store (one instruction per box); MODE 1: 1-D bulk stores, one per row and level, spread over the threads
template <int TX, int TY, int LZ, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) k_tma(const __grid_constant__ CUtensorMap map, float* out, int ox, int oy, int nz, int chunk)
{
    extern __shared__ __align__(128) float s_out[]; // [2][LZ][TY][TX]
    constexpr int kBox = TX * TY * LZ;
    const int tiles_x = (ox + TX - 1) / TX;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int z0 = blockIdx.y * chunk, z1 = min(nz, z0 + chunk);
    int buf = 0;
    for (int z = z0; z < z1; z += LZ, buf ^= 1) {
        float* s = s_out + buf * kBox;
        if (threadIdx.x == 0)
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); // the store that read this buffer two batches ago is done with it
        __syncthreads();
#pragma unroll 8
        for (int i = threadIdx.x; i < kBox; i += THREADS)
            s[i] = (float)(z + i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (MODE == 0) {
            if (threadIdx.x == 0) {
                tma_store_3d(&map, s, tx * TX, ty * TY, z);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            // rows of the box: LZ * TY of them, TX floats each
            for (int r = threadIdx.x; r < LZ * TY; r += THREADS) {
                const int lz = r / TY, ly = r % TY;
                if (z + lz < z1 && ty * TY + ly < oy)
                    bulk_store_1d(out + ((size_t)(z + lz) * oy + (ty * TY + ly)) * ox + tx * TX, s + r * TX, TX * 4);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // (MODE 1 waits per thread: every thread that issued copies tracks its own groups)
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// reference pattern: per-thread STG as the shipped kernels do it
template <int TX, int TY, int VEC>
__global__ void k_tiles(float* out, int ox, int oy, int nz, int chunk)
{
    const int tiles_x = ox / TX;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    if (tx >= tiles_x) return;
    const int z0 = blockIdx.y * chunk, z1 = min(nz, z0 + chunk);
    constexpr int per_row = TX / VEC;
    constexpr int rows_per_pass = 256 / per_row;
    const int lx = (threadIdx.x % per_row) * VEC, ly = threadIdx.x / per_row;
    const size_t level = (size_t)ox * oy;
    for (int z = z0; z < z1; ++z) {
        float* base = out + z * level + (size_t)(ty * TY) * ox + tx * TX + lx;
#pragma unroll
        for (int r = ly; r < TY; r += rows_per_pass) {
            if (VEC == 4)
                __stcs(reinterpret_cast<float4*>(base + (size_t)r * ox), make_float4(1.f, 2.f, 3.f, (float)z));
            else
                __stcs(base + (size_t)r * ox, (float)z);
        }
    }
}

template <class F>
float time_ms(F f, int reps = 5)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

template <int TX, int TY, int LZ, int MODE, int THREADS = 256>
void run_tma(float* out, int ox, int oy, int nz, int chunk, const char* what)
{
    CUtensorMap map = make_map(out, ox, oy, nz, TX, TY, LZ);
    const int smem = 2 * TX * TY * LZ * 4;
    CK(cudaFuncSetAttribute(k_tma<TX, TY, LZ, MODE, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 g(((ox + TX - 1) / TX) * ((oy + TY - 1) / TY), (nz + chunk - 1) / chunk);
    const double gb = (double)ox * oy * nz * 4 / 1e9;
    float ms = time_ms([&] { k_tma<TX, TY, LZ, MODE, THREADS><<<g, THREADS, smem>>>(map, out, ox, oy, nz, chunk); });
    CK(cudaGetLastError());
    printf("%-9s box %3dx%2dx%d (%3d KB smem, %d thr) chunk %3d %8.3f ms %8.1f GB/s\n", what, TX, TY, LZ, smem / 1024, THREADS, chunk, ms, gb / ms * 1e3);
}

// ---- LDS.128 with part of the lanes active --------------------------------------------------------------------------------
template <int W>
__global__ void k_lds_mask(float* out, int iters, long long* cyc, unsigned mask)
{
    __shared__ __align__(16) float s[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int idx = ((lane * 12 + w * 16) & 4095) / 4 * 4;
    const bool on = (mask >> lane) & 1u;
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int a = (idx + u * 512) & 8191;
            if (on) {
                if (W == 4) { float4 v = *reinterpret_cast<const float4*>(s + a); acc += v.x + v.y + v.z + v.w; }
                else acc += s[a];
            }
        }
        idx = ((idx + (int)(acc == 12345.f)) & 4095) / 4 * 4;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int W>
void run_lds(const char* name, unsigned mask)
{
    float* out; long long* cyc; CK(cudaMalloc(&out, 148 * 1024 * 4)); CK(cudaMalloc(&cyc, 148 * 8));
    const int iters = 4000, threads = 512;
    k_lds_mask<W><<<148, threads>>>(out, iters, cyc, mask);
    k_lds_mask<W><<<148, threads>>>(out, iters, cyc, mask);
    CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("%-44s mask %08x  %6.2f cycles per warp load\n", name, mask, (double)h[0] / (iters * 8.0 * (threads / 32)));
    cudaFree(out); cudaFree(cyc);
}

int main(int argc, char** argv)
{
    const int ox = argc > 1 ? atoi(argv[1]) : 2000, oy = 2000, nz = 1644;
    printf("ox = %d (row pitch %d B), %d x %d x %d floats\n", ox, 4 * ox, ox, oy, nz);
    const size_t n = (size_t)ox * oy * nz;
    float* out; CK(cudaMalloc(&out, n * 4));
    const double gb = n * 4 / 1e9;
    float ms = time_ms([&] { CK(cudaMemsetAsync(out, 0, n * 4)); });
    printf("cudaMemset                  %8.3f ms %8.1f GB/s\n", ms, gb / ms * 1e3);
    for (int chunk : {64, 128}) {
        { dim3 g((ox / 64) * (oy / 16), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles<64, 16, 1><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("STG  tiles 64x16 scalar chunk %3d %8.3f ms %8.1f GB/s (of %d columns)\n", chunk, ms, 4e-9 * (ox / 64 * 64) * (double)oy * nz / ms * 1e3, ox / 64 * 64); }
        { dim3 g((ox / 128) * (oy / 8), (nz + chunk - 1) / chunk);
          ms = time_ms([&] { k_tiles<128, 8, 4><<<g, 256>>>(out, ox, oy, nz, chunk); });
          printf("STG  tiles 128x8 float4 chunk %3d %8.3f ms %8.1f GB/s (of %d columns)\n", chunk, ms, 4e-9 * (ox / 128 * 128) * (double)oy * nz / ms * 1e3, ox / 128 * 128); }
        run_tma<64, 16, 8, 0>(out, ox, oy, nz, chunk, "TMA 3-D");
        run_tma<64, 16, 4, 0>(out, ox, oy, nz, chunk, "TMA 3-D");
        run_tma<128, 8, 8, 0>(out, ox, oy, nz, chunk, "TMA 3-D");
        run_tma<128, 8, 4, 0>(out, ox, oy, nz, chunk, "TMA 3-D");
        run_tma<256, 4, 8, 0>(out, ox, oy, nz, chunk, "TMA 3-D");
        run_tma<256, 4, 4, 0>(out, ox, oy, nz, chunk, "TMA 3-D");
        run_tma<256, 8, 4, 0>(out, ox, oy, nz, chunk, "TMA 3-D");
        run_tma<128, 16, 4, 0>(out, ox, oy, nz, chunk, "TMA 3-D");
        run_tma<256, 16, 2, 0, 512>(out, ox, oy, nz, chunk, "TMA 3-D");
        run_tma<32, 32, 8, 0>(out, ox, oy, nz, chunk, "TMA 3-D");
        run_tma<64, 16, 8, 1>(out, ox, oy, nz, chunk, "bulk 1-D");
        run_tma<128, 8, 8, 1>(out, ox, oy, nz, chunk, "bulk 1-D");
        run_tma<256, 4, 8, 1>(out, ox, oy, nz, chunk, "bulk 1-D");
    }
    printf("\nshared loads with part of the lanes active (32 distinct 16-byte-aligned addresses, stride 12 words)\n");
    run_lds<4>("LDS.128 all lanes", 0xffffffffu);
    run_lds<4>("LDS.128 16 lanes (every other)", 0x55555555u);
    run_lds<4>("LDS.128 16 lanes (lower half)", 0x0000ffffu);
    run_lds<4>("LDS.128 8 lanes (every 4th)", 0x11111111u);
    run_lds<4>("LDS.128 8 lanes (one quarter)", 0x000000ffu);
    run_lds<4>("LDS.128 4 lanes (every 8th)", 0x01010101u);
    run_lds<4>("LDS.128 2 lanes", 0x00010001u);
    run_lds<1>("LDS.32  all lanes", 0xffffffffu);
    run_lds<1>("LDS.32  8 lanes (every 4th)", 0x11111111u);
    return 0;
}
