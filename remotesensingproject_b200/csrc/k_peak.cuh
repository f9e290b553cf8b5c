/*
 * k_peak.cuh — on-device measurement of the FP32 (non-tensor) issue rate, the
 * roofline denominator of the mean-shift kernel.  Two figures:
 *   nofma: separately rounded FMUL / FADD in the 1:1 mix the parity rules force
 *          on the hot loop (one operation per instruction);
 *   fma  : FFMA, counted as 2 flops (the usual "peak FP32" figure).
 */
#pragma once
#include "rslf_common.cuh"

template <bool FMA>
__global__ void __launch_bounds__(256)
fp32_peak_kernel(float* out, int iters, float a, float b)
{
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (FMA) x[i] = __fmaf_rn(x[i], a, b);
            else { x[i] = __fmul_rn(x[i], a); x[i] = __fadd_rn(x[i], b); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 12345.678f) out[0] = s;        /* keeps the chain alive; practically never true */
}

/* packed FP32x2 (Blackwell FFMA2 / FADD2), separately rounded: mul as fma(a, b, -0), then add */
__device__ __forceinline__ unsigned long long mul2_rn(unsigned long long a, unsigned long long b, unsigned long long nz) {
    unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz)); return r;
}
__device__ __forceinline__ unsigned long long add2_rn(unsigned long long a, unsigned long long b) {
    unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__global__ void __launch_bounds__(256)
fp32x2_peak_kernel(float* out, int iters, float a, float b)
{
    unsigned long long x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = pk2((float)(threadIdx.x + i) * 1e-3f, (float)(threadIdx.x + i) * 2e-3f);
    const unsigned long long A = pk2(a, a), B = pk2(b, b), NZ = pk2(-0.f, -0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = mul2_rn(x[i], A, NZ); x[i] = add2_rn(x[i], B); }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float p, q; asm("mov.b64 {%0,%1}, %2;" : "=f"(p), "=f"(q) : "l"(x[i])); s += p + q; }
    if (s == 12345.678f) out[0] = s;
}

static int measure_fp32x2_peak(rslf_ctx* ctx, double* gops)
{
    float* sink = nullptr;
    RSLF_CUDA_TRY(ctx, cudaMalloc((void**)&sink, 4));
    const int blocks = ctx->num_sm * 8, threads = 256, iters = 8192;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, ctx->stream);
        fp32x2_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(sink, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { cudaFree(sink); RSLF_CUDA_TRY(ctx, e); }
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        /* 16 packed instructions = 32 separately rounded operations per inner iteration per thread */
        double rate = (double)blocks * threads * iters * 32.0 / (ms * 1e-3) * 1e-9;
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    if (gops) *gops = best;
    return RSLF_OK;
}

static int measure_fp32_peak(rslf_ctx* ctx, double* gops_nofma, double* gflops_fma)
{
    float* sink = nullptr;
    RSLF_CUDA_TRY(ctx, cudaMalloc((void**)&sink, 4));
    const int blocks = ctx->num_sm * 8, threads = 256, iters = 8192;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best[2] = {0, 0};
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0, ctx->stream);
            if (mode == 0) fp32_peak_kernel<false><<<blocks, threads, 0, ctx->stream>>>(sink, iters, 0.999f, 1e-3f);
            else fp32_peak_kernel<true><<<blocks, threads, 0, ctx->stream>>>(sink, iters, 0.999f, 1e-3f);
            cudaEventRecord(e1, ctx->stream);
            cudaError_t e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { cudaFree(sink); RSLF_CUDA_TRY(ctx, e); }
            float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
            /* instructions: nofma = 32 per inner iteration per thread, fma = 16 (x2 flops) */
            double ops = (double)blocks * threads * iters * 32.0;
            double rate = ops / (ms * 1e-3) * 1e-9;
            if (rep > 0 && rate > best[mode]) best[mode] = rate;
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    if (gops_nofma) *gops_nofma = best[0];
    if (gflops_fma) *gflops_fma = best[1];
    return RSLF_OK;
}
