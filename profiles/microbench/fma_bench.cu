// Microbenchmark: scalar FFMA vs packed FFMA2 throughput on sm_100a (per-SM FMA rate), various warps/SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rd) : "f"(d.x), "f"(d.y));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rd) : "l"(ra), "l"(rb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
}
__device__ __forceinline__ void ffma2s(float2& d, float a, float2 b) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ra) : "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rd) : "f"(d.x), "f"(d.y));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rd) : "l"(ra), "l"(rb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
}
template <int MODE>
__global__ void k(float* out, int iters, float a0, float b0) {
    float2 acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float2 a = make_float2(a0, a0 * 1.0001f), b = make_float2(b0, b0 * 0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a.x, b.x); acc[i].y = fmaf(acc[i].y, a.y, b.y); }
                else if (MODE == 1) { float2 t = b; ffma2(t, acc[i], a); acc[i] = t; }
                else if (MODE == 2) { ffma2s(acc[i], a.x + i, b); }   // d += scalar * vec (GEMM form)
                else if (MODE == 3) { ffma2(acc[i], a, b); }          // d += vec * vec
            }
        }
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) r += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 64 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    for (int threads : {128, 256, 512, 1024}) {
        for (int mode = 0; mode < 4; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<148, threads>>>(out, iters, 1.0001f, 0.5f);
                else if (mode == 1) k<1><<<148, threads>>>(out, iters, 1.0001f, 0.5f);
                else if (mode == 2) k<2><<<148, threads>>>(out, iters, 1.0001f, 0.5f);
                else k<3><<<148, threads>>>(out, iters, 1.0001f, 0.5f);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                double fma = 148.0 * threads * iters * 4 * 8 * 2;
                if (rep) printf("threads/SM %4d  %s : %.3f ms  %.1f TFLOP/s  (%.1f FMA/clk/SM @1.965GHz)\n", threads, mode == 0 ? "FFMA " : mode == 1 ? "FFMA2 (d=a*b+c)" : mode == 2 ? "FFMA2 acc+=s*v" : "FFMA2 acc+=v*v", ms, 2 * fma / ms / 1e9, fma / 148 / (ms * 1e-3) / 1.965e9);
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
