// Microbenchmark: cycles per LDS.128 for different lane-address patterns (1 warp per SM and 8 warps per SM).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int pattern, int iters, long long* out, float* sink) {
    __shared__ __align__(16) float buf[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = i * 0.001f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int off;  // in floats, multiple of 4
    switch (pattern) {
        case 0: off = 0; break;                                   // all lanes same 16 B
        case 1: off = (lane >> 3) * 4; break;                     // one address per quarter-warp, 4 distinct (64 B)
        case 2: off = (lane & 7) * 4; break;                      // 8 distinct (128 B) per quarter, same for all quarters
        case 3: off = lane * 4; break;                            // 32 distinct, 512 B contiguous
        case 4: off = (lane >> 3) * 4 * 33; break;                // one address per quarter, far apart rows (stride 132 B)
        case 5: off = (lane & 7) * 132; break;                    // 8 rows (stride 528 B), same for all quarters
        case 6: off = (lane >> 4) * 4; break;                     // one address per half-warp
        case 7: off = (lane & 15) * 4; break;                     // 16 distinct per half, same both halves
        case 8: off = (lane >> 2) * 4; break;                     // 8 distinct groups of 4 lanes (v2 weight pattern: og = lane>>2)
        case 9: off = (lane & 3) * 4; break;                      // 4 distinct (v2 act pattern: sg = lane&3)
        default: off = 0;
    }
    float4 acc = make_float4(0, 0, 0, 0);
    const float* p = buf + off;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            float4 v;
            unsigned a = (unsigned)__cvta_generic_to_shared(p + ((it * 16 + u) & 7) * 1024);
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
            if (u & 1) acc.x += v.x; else acc.y += v.w;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc.x + acc.y + acc.z + acc.w == 12345.f) sink[0] = acc.x;
}
int main() {
    long long* out; float* sink; cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
    const int iters = 2000;
    const char* names[] = {"all lanes same 16B", "1 addr per quarter (4 distinct, 64B)", "8 distinct per quarter, quarters same (128B)",
                           "32 distinct (512B)", "1 addr/quarter, rows 132B apart", "8 rows stride 528B, quarters same",
                           "1 addr per half-warp", "16 distinct per half, halves same (256B)", "8 groups of 4 lanes (128B)", "4 distinct, lane&3 (64B)"};
    for (int threads : {32, 256, 1024}) {
        for (int pat = 0; pat < 10; ++pat) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) { k<<<148, threads>>>(pat, iters, out, sink); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost); }
            printf("warps/SM %d  pattern %d %-48s : %.2f cycles per LDS.128 per warp, %.2f SM-cycles per warp-LDS\n", threads / 32, pat, names[pat],
                   (double)h / (iters * 16.0), (double)h / (iters * 16.0) / (threads / 32));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
