// Probe (GPU box): can TMEM (tcgen05.st / tcgen05.ld) serve as per-warp radiance storage for the mean shift?
// Each warp repeatedly reads 12 floats per lane (= one 4-view RGB block) and runs 80 separately rounded FP32
// operations on them (the mean-shift budget of 4 views), from (a) nothing (registers only), (b) shared memory
// (3 x LDS.128), (c) tensor memory (3 x tcgen05.ld.32x32b.x4).  Prints cycles per block and warp at 4, 8, 12 resident
// warps per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tmem_probe tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float& a, float& b, float& c, float& d)
{
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
    a = __uint_as_float(r0); b = __uint_as_float(r1); c = __uint_as_float(r2); d = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)),
                 "r"(__float_as_uint(c)), "r"(__float_as_uint(d)) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 4 views x 3 channels, mean-shift arithmetic (19 FP32 + 1 max per view)
__device__ __forceinline__ void work(const float (&r)[3][4], const float (&rb)[3], float inv, float (&sR)[3], float& sK)
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float x0 = r[0][j] - rb[0], x1 = r[1][j] - rb[1], x2 = r[2][j] - rb[2];
        const float b0 = (inv * x0) * x0, b1 = (inv * x1) * x1, b2 = (inv * x2) * x2;
        const float b = (b0 + b1) + b2;
        const float k = fmaxf(1.0f - b, 0.f);
        sR[0] = sR[0] + r[0][j] * k; sR[1] = sR[1] + r[1][j] * k; sR[2] = sR[2] + r[2][j] * k;
        sK = sK + k;
    }
}

template <int MODE>   // 0 registers, 1 shared memory, 2 tensor memory
__global__ void __launch_bounds__(384, 1) probe(float* out, long long* cyc, int iters, int nblocks_views, float inv)
{
    extern __shared__ float4 smem[];
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int COLS = 512;
    if (MODE == 2) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "n"(COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    const uint32_t tbase = (MODE == 2) ? tmem_base_s + ((uint32_t)(32 * (warp & 3)) << 16) + 128u * (uint32_t)(warp >> 2) : 0;
    float* mine = reinterpret_cast<float*>(smem) + warp * (nblocks_views * 12 * 32);
    // fill
    for (int b = 0; b < nblocks_views; ++b)
        for (int c = 0; c < 3; ++c) {
            const float v0 = 0.001f * (lane + b + c), v1 = v0 + 0.01f, v2 = v0 + 0.02f, v3 = v0 + 0.03f;
            if (MODE == 1) *reinterpret_cast<float4*>(mine + ((b * 3 + c) * 32 + lane) * 4) = make_float4(v0, v1, v2, v3);
            if (MODE == 2) tmem_st4(tbase + (b * 12 + c * 4), v0, v1, v2, v3);
        }
    if (MODE == 2) tmem_wait_st();
    __syncwarp();
    float rb[3] = {0.1f, 0.2f, 0.3f}, sR[3] = {0, 0, 0}, sK = 0;
    float ra[3][4], rbuf[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) { ra[c][j] = 0.01f * (c + j + lane); rbuf[c][j] = ra[c][j] + 0.5f; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        // ping-pong over the blocks like the depth kernel: load next, compute current
        auto load = [&](int b, float (&dst)[3][4]) {
            if (MODE == 1) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float4 t = *reinterpret_cast<const float4*>(mine + ((b * 3 + c) * 32 + lane) * 4);
                    dst[c][0] = t.x; dst[c][1] = t.y; dst[c][2] = t.z; dst[c][3] = t.w;
                }
            } else if (MODE == 2) {
#pragma unroll
                for (int c = 0; c < 3; ++c) tmem_ld4(tbase + (b * 12 + c * 4), dst[c][0], dst[c][1], dst[c][2], dst[c][3]);
            }
        };
        load(0, ra);
        if (MODE == 2) tmem_wait_ld();
        for (int b = 0; b + 2 <= nblocks_views; b += 2) {
            load(b + 1, rbuf);
            work(ra, rb, inv, sR, sK);
            if (MODE == 2) tmem_wait_ld();
            load(min(b + 2, nblocks_views - 1), ra);
            work(rbuf, rb, inv, sR, sK);
            if (MODE == 2) tmem_wait_ld();
        }
        rb[0] = sR[0] * 1e-3f; rb[1] = sR[1] * 1e-3f; rb[2] = sR[2] * 1e-3f;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sR[0] + sR[1] + sR[2] + sK;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (MODE == 2) {
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "n"(COLS));
    }
}

template <int MODE>
static int run(const char* name, int warps_per_cta, int sms)
{
    const int nblk = 10, iters = 2000;            // 10 blocks of 4 views = 120 columns per warp
    const int threads = 32 * warps_per_cta;
    const size_t dyn = (MODE == 1) ? (size_t)warps_per_cta * nblk * 12 * 32 * 4 : 0;
    CHECK(cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(dyn > 0 ? dyn : 1024)));
    int occ = 0;
    CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe<MODE>, threads, dyn));
    const int grid = sms;                           // one CTA per SM
    float* out; long long* cyc;
    CHECK(cudaMalloc(&out, (size_t)grid * threads * 4));
    CHECK(cudaMalloc(&cyc, (size_t)grid * 8));
    probe<MODE><<<grid, threads, dyn>>>(out, cyc, 10, nblk, 25.f);
    CHECK(cudaDeviceSynchronize());
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    probe<MODE><<<grid, threads, dyn>>>(out, cyc, iters, nblk, 25.f);
    cudaEventRecord(b);
    CHECK(cudaDeviceSynchronize());
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    long long h = 0; CHECK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    const double blocks = (double)iters * nblk;                       // 4-view blocks per warp
    const double warps = (double)grid * warps_per_cta;
    printf("%-8s %2d warps/SM (1 CTA, occupancy API says %d): %8.3f ms, %7.1f cycles per 4-view block and warp (issue floor %d), %.3e view-samples/s\n",
           name, warps_per_cta, occ, ms, (double)h / blocks, 86 * warps_per_cta / 4, warps * blocks * 4 * 32 / (ms * 1e-3));
    cudaFree(out); cudaFree(cyc);
    return 0;
}

int main()
{
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    for (int w = 4; w <= 12; w += 4) {
        if (run<1>("shared", w, sms)) return 1;
        if (run<2>("tmem", w, sms)) return 1;
    }
    return 0;
}
