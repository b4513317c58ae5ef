// Probe for the shared-memory descriptor conventions the tensor-core training kernel relies on (sm_100a):
//   T1  D[s][j]  = sum_k A[s][k]  W[j][k]    A K-major panels, B K-major panels            (forward layer)
//   T2  D[s][k]  = sum_j dZ[s][j] W[j][k]    A K-major panels, B = the SAME W panels read MN-major  (dh)
//   T2b same with N = 16
//   T3  D[k][j]  = sum_s h[s][k]  dZ[s][j]   A = activation panels read MN-major (M = 128 spans 32 panels),
//                                            B = dZ panels read MN-major, K = 128 samples in 16 steps   (dW)
//   T4  tcgen05.st / tcgen05.ld round trip from warps 4..7 (lane quarter = warp % 4) on a second column range
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o dbg/umma_mn dbg/umma_mn.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int TILE = 128, HP = 32;
constexpr int kPanelA = TILE * 16, kPanelB = HP * 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, int amn, int bmn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::
            "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// smem: Hs = activation panels (9 panels: 32 features + a ones panel), Ds = dZ panels (8), Ws = weight panels (8)
// inputs (global): h[128][32], dz[128][32], W[32][32]; outputs: D1[128][32], D2[128][32], D2b[128][16], D3[128][32], D4[256][16]
__global__ void __launch_bounds__(256, 1) probe(const float* h, const float* dz, const float* W, float* D1, float* D2,
                                                float* D2b, float* D3, float* D4) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* Hs = smem + 128;                 // 32 panels of room (64 KB) so the M = 128 MN-major read stays in bounds
    unsigned char* Ds = Hs + 32 * kPanelA;
    unsigned char* Ws = Ds + 8 * kPanelA;
    const uint32_t bar = smem_u32(smem);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 16);
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (t < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    for (int e = t; e < 32 * kPanelA / 4; e += blockDim.x) reinterpret_cast<float*>(Hs)[e] = 0.0f;
    __syncthreads();
    if (t < TILE) {
        for (int k = 0; k < 32; ++k) {
            *reinterpret_cast<float*>(Hs + (k >> 2) * kPanelA + t * 16 + (k & 3) * 4) = h[t * 32 + k];
            *reinterpret_cast<float*>(Ds + (k >> 2) * kPanelA + t * 16 + (k & 3) * 4) = dz[t * 32 + k];
        }
        *reinterpret_cast<float*>(Hs + 8 * kPanelA + t * 16) = 1.0f;  // ones column = feature 32
    }
    for (int e = t; e < 32 * 32; e += blockDim.x) {
        const int j = e / 32, k = e % 32;
        *reinterpret_cast<float*>(Ws + (k >> 2) * kPanelB + j * 16 + (k & 3) * 4) = W[e];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *slot;
    const uint32_t aH = smem_u32(Hs), aD = smem_u32(Ds), aW = smem_u32(Ws);
    if (t == 0) {
        // T1 -> columns [0,32)
        for (int ks = 0; ks < 4; ++ks)
            mma(tm + 0, make_desc(aH + ks * 2 * kPanelA, kPanelA, 128), make_desc(aW + ks * 2 * kPanelB, kPanelB, 128),
                idesc(128, 32, 0, 0), ks > 0);
        // T2 -> columns [32,64): B MN-major over the W panels: N chunk stride (SBO) = kPanelB, 8 j rows of 16 B, K step = 128 B
        for (int ks = 0; ks < 4; ++ks)
            mma(tm + 32, make_desc(aD + ks * 2 * kPanelA, kPanelA, 128), make_desc(aW + ks * 128, 128, kPanelB),
                idesc(128, 32, 0, 1), ks > 0);
        // T2b -> columns [64,80), N = 16
        for (int ks = 0; ks < 4; ++ks)
            mma(tm + 64, make_desc(aD + ks * 2 * kPanelA, kPanelA, 128), make_desc(aW + ks * 128, 128, kPanelB),
                idesc(128, 16, 0, 1), ks > 0);
        // T3 -> columns [96,128): A MN-major over the activation panels (SBO = kPanelA), B MN-major over dZ panels
        for (int ks = 0; ks < 16; ++ks)
            mma(tm + 96, make_desc(aH + ks * 128, 128, kPanelA), make_desc(aD + ks * 128, 128, kPanelA),
                idesc(128, 32, 1, 1), ks > 0);
        commit(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32) << 16);
    if (t < TILE) {
        uint32_t r[32];
        ld32(tm + lane_base + 0, r);
        for (int c = 0; c < 32; ++c) D1[t * 32 + c] = __uint_as_float(r[c]);
        ld32(tm + lane_base + 32, r);
        for (int c = 0; c < 32; ++c) D2[t * 32 + c] = __uint_as_float(r[c]);
        ld32(tm + lane_base + 64, r);
        for (int c = 0; c < 16; ++c) D2b[t * 16 + c] = __uint_as_float(r[c]);
        ld32(tm + lane_base + 96, r);
        for (int c = 0; c < 32; ++c) D3[t * 32 + c] = __uint_as_float(r[c]);
    }
    // T4: every thread (8 warps) stores 16 values into columns [128 + 16*(warp/4), +16) of its lane, reads them back
    {
        uint32_t v[16], w[16];
        for (int c = 0; c < 16; ++c) v[c] = __float_as_uint((float)(t * 100 + c));
        const uint32_t addr = tm + lane_base + 128 + 16 * (warp >> 2);
        st16(addr, v);
        ld16(addr, w);
        for (int c = 0; c < 16; ++c) D4[t * 16 + c] = __uint_as_float(w[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm));
}

int main() {
    const int nh = 128 * 32;
    float *h = new float[nh], *dz = new float[nh], *W = new float[1024];
    srand(1);
    for (int i = 0; i < nh; ++i) { h[i] = (float)(rand() % 17 - 8); dz[i] = (float)(rand() % 13 - 6); }
    for (int i = 0; i < 1024; ++i) W[i] = (float)(rand() % 9 - 4);
    float *dh, *ddz, *dW, *d1, *d2, *d2b, *d3, *d4;
    cudaMalloc(&dh, nh * 4); cudaMalloc(&ddz, nh * 4); cudaMalloc(&dW, 4096);
    cudaMalloc(&d1, nh * 4); cudaMalloc(&d2, nh * 4); cudaMalloc(&d2b, nh * 2); cudaMalloc(&d3, nh * 4); cudaMalloc(&d4, 256 * 16 * 4);
    cudaMemcpy(dh, h, nh * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(ddz, dz, nh * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W, 4096, cudaMemcpyHostToDevice);
    const int smem = 128 + 32 * kPanelA + 8 * kPanelA + 8 * kPanelB;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 256, smem>>>(dh, ddz, dW, d1, d2, d2b, d3, d4);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    float *o1 = new float[nh], *o2 = new float[nh], *o2b = new float[nh / 2], *o3 = new float[nh], *o4 = new float[4096];
    cudaMemcpy(o1, d1, nh * 4, cudaMemcpyDeviceToHost); cudaMemcpy(o2, d2, nh * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(o2b, d2b, nh * 2, cudaMemcpyDeviceToHost); cudaMemcpy(o3, d3, nh * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(o4, d4, 4096 * 4, cudaMemcpyDeviceToHost);
    int bad1 = 0, bad2 = 0, bad2b = 0, bad3 = 0, bad3b = 0, bad4 = 0;
    for (int s = 0; s < 128; ++s)
        for (int j = 0; j < 32; ++j) {
            float r1 = 0, r2 = 0;
            for (int k = 0; k < 32; ++k) { r1 += h[s * 32 + k] * W[j * 32 + k]; r2 += dz[s * 32 + k] * W[k * 32 + j]; }
            if (o1[s * 32 + j] != r1) ++bad1;
            if (o2[s * 32 + j] != r2) ++bad2;
            if (j < 16 && o2b[s * 16 + j] != r2) ++bad2b;
        }
    for (int k = 0; k < 33; ++k)
        for (int j = 0; j < 32; ++j) {
            float r = 0;
            for (int s = 0; s < 128; ++s) r += (k < 32 ? h[s * 32 + k] : 1.0f) * dz[s * 32 + j];
            if (o3[k * 32 + j] != r) { if (k < 32) ++bad3; else ++bad3b; }
        }
    for (int t = 0; t < 256; ++t)
        for (int c = 0; c < 16; ++c) if (o4[t * 16 + c] != (float)(t * 100 + c)) ++bad4;
    printf("T1 fwd (K,K) mismatches: %d\nT2 dh (K,MN) mismatches: %d\nT2b dh N=16 mismatches: %d\n"
           "T3 dW (MN,MN) mismatches: %d, bias row mismatches: %d\nT4 st/ld round trip mismatches: %d\n",
           bad1, bad2, bad2b, bad3, bad3b, bad4);
    if (bad2) { printf("T2 sample row0: "); for (int j = 0; j < 8; ++j) printf("%g ", o2[j]); printf("\n"); }
    if (bad3) { printf("T3 sample row0: "); for (int j = 0; j < 8; ++j) printf("%g ", o3[j]); printf("\n"); }
    return 0;
}
