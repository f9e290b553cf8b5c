// ffma2_probe.cu — does the packed FP32x2 pipe (FFMA2 / FADD2) speed up the CONTRACTED mean shift?
// Measures, per lane, view-iterations per second of the mean-shift body on register-resident radiances:
//   mode 0: scalar FMA form (ms_accumulate<FAST>, 12 instructions per view)
//   mode 1: packed form, two views per instruction (11 packed + 2 FMNMX per two views)
//   mode 2: packed FFMA2 with three fresh operand pairs (raw pipe rate)   mode 3: scalar FFMA, three fresh operands
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ffma2_probe ffma2_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

constexpr int NV = 16;   // views held in registers (the real kernel streams 100 from smem / TMEM)

template <int MODE>
__global__ void __launch_bounds__(384, 1) probe(const float* __restrict__ in, float* __restrict__ out, int iters, float inv)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float r[NV][3];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int c = 0; c < 3; ++c) r[v][c] = in[(tid * NV + v) * 3 + c];
    float rb0 = r[0][0], rb1 = r[0][1], rb2 = r[0][2];
    float acc = 0.f;
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, sk = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float x0 = r[v][0] - rb0, x1 = r[v][1] - rb1, x2 = r[v][2] - rb2;
                float t = x0 * x0; t = __fmaf_rn(x1, x1, t); t = __fmaf_rn(x2, x2, t);
                const float k = fmaxf(__fmaf_rn(-inv, t, 1.f), 0.f);
                s0 = __fmaf_rn(r[v][0], k, s0); s1 = __fmaf_rn(r[v][1], k, s1); s2 = __fmaf_rn(r[v][2], k, s2);
                sk += k;
            }
            const float q = __frcp_rn(sk + 1e-20f);
            rb0 = s0 * q; rb1 = s1 * q; rb2 = s2 * q; acc += sk;
        }
    } else if (MODE == 1) {
        f32x2 R[NV / 2][3];
#pragma unroll
        for (int v = 0; v < NV / 2; ++v)
#pragma unroll
            for (int c = 0; c < 3; ++c) R[v][c] = pk2(r[2 * v][c], r[2 * v + 1][c]);
        const f32x2 NINV = pk2(-inv, -inv), ONE = pk2(1.f, 1.f);
        for (int it = 0; it < iters; ++it) {
            f32x2 S0 = pk2(0.f, 0.f), S1 = S0, S2 = S0, SK = S0;
            const f32x2 B0 = pk2(rb0, rb0), B1 = pk2(rb1, rb1), B2 = pk2(rb2, rb2);
#pragma unroll
            for (int v = 0; v < NV / 2; ++v) {
                const f32x2 X0 = sub2(R[v][0], B0), X1 = sub2(R[v][1], B1), X2 = sub2(R[v][2], B2);
                f32x2 T = mul2(X0, X0); T = fma2(X1, X1, T); T = fma2(X2, X2, T);
                float ka, kb;
                upk2(fma2(NINV, T, ONE), ka, kb);
                const f32x2 K = pk2(fmaxf(ka, 0.f), fmaxf(kb, 0.f));
                S0 = fma2(R[v][0], K, S0); S1 = fma2(R[v][1], K, S1); S2 = fma2(R[v][2], K, S2);
                SK = add2(SK, K);
            }
            float a, b; upk2(SK, a, b); const float sk = a + b;
            const float q = __frcp_rn(sk + 1e-20f);
            upk2(S0, a, b); rb0 = (a + b) * q; upk2(S1, a, b); rb1 = (a + b) * q; upk2(S2, a, b); rb2 = (a + b) * q;
            acc += sk;
        }
    } else if (MODE == 2) {
        f32x2 A[8], B[8], Cc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { A[i] = pk2(r[i][0], r[i][1]); B[i] = pk2(r[i][2], r[i + 8][0]); Cc[i] = pk2(r[i + 8][1], r[i + 8][2]); }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 3; ++rep)
#pragma unroll
                for (int i = 0; i < 8; ++i) Cc[i] = fma2(A[i], B[(i + rep) & 7], Cc[i]);
        }
        float a, b;
#pragma unroll
        for (int i = 0; i < 8; ++i) { upk2(Cc[i], a, b); acc += a + b; }
    } else {
        float A[16], B[16], Cc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { A[i] = r[i][0]; B[i] = r[i][1]; Cc[i] = r[i][2]; }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 3; ++rep)
#pragma unroll
                for (int i = 0; i < 16; ++i) Cc[i] = __fmaf_rn(A[i], B[(i + rep) & 15], Cc[i]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += Cc[i];
    }
    out[tid] = acc + rb0 + rb1 + rb2;
}

template <int MODE> double run(const float* in, float* out, int iters, int sms)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe<MODE><<<sms, 384>>>(in, out, 10, 25.f);
    cudaEventRecord(a);
    probe<MODE><<<sms, 384>>>(in, out, iters, 25.f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    if (cudaGetLastError() != cudaSuccess) { printf("mode %d failed\n", MODE); return 0; }
    return ms;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, threads = sms * 384, iters = 20000;
    float *in, *out; cudaMalloc(&in, (size_t)threads * NV * 3 * 4); cudaMalloc(&out, threads * 4);
    float* h = (float*)malloc((size_t)threads * NV * 3 * 4);
    for (size_t i = 0; i < (size_t)threads * NV * 3; ++i) h[i] = 0.3f + 0.1f * (float)(rand() % 1000) / 1000.f;
    cudaMemcpy(in, h, (size_t)threads * NV * 3 * 4, cudaMemcpyHostToDevice);
    const double m0 = run<0>(in, out, iters, sms), m1 = run<1>(in, out, iters, sms), m2 = run<2>(in, out, iters, sms), m3 = run<3>(in, out, iters, sms);
    const double vi = (double)threads * iters * NV;
    printf("{\"sms\": %d, \"scalar_fma_meanshift_Gviewiter_s\": %.1f, \"packed_meanshift_Gviewiter_s\": %.1f, \"packed_speedup\": %.3f, "
           "\"ffma2_fresh_operands_Gflop_s\": %.1f, \"ffma_fresh_operands_Gflop_s\": %.1f}\n",
           sms, vi / m0 * 1e-6, vi / m1 * 1e-6, m0 / m1, (double)threads * iters * 24 * 4 / m2 * 1e-6, (double)threads * iters * 48 * 2 / m3 * 1e-6);
    return 0;
}
