// Empirical address map of tcgen05.mma shared-memory operands (kind::tf32): one operand region is filled with its own
// word indices, the other operand is a K-major identity, so the accumulator shows which word the hardware reads for
// every (row, k).  build: nvcc -gencode arch=compute_100a,code=sm_100a -o dbg/umma_addr dbg/umma_addr.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7u) << 61;
    return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc)
        : "memory");
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

struct Cfg { uint32_t idesc, which /*0: probe B, 1: probe A*/, lbo, sbo, layout, kind /*0 tf32, 1 bf16*/, pad; };

// region X (16 KB = 4096 words, value = word index & 2047, plus a tag in the high words) at smem+1024;
// identity operand (K-major, 128 rows x 8: I[r][k] = (r == k)) at smem + 1024 + 65536
__global__ void __launch_bounds__(128, 1) probe(Cfg c, float* D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* X = reinterpret_cast<float*>(smem + 1024);
    unsigned char* I = smem + 1024 + 65536;
    const uint32_t bar = smem_u32(smem);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 16);
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (t < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (c.kind == 0) {
        for (int e = t; e < 16384; e += blockDim.x) X[e] = (float)(e & 2047);   // 64 KB, value repeats every 8 KB
    } else {
        unsigned short* Xh = reinterpret_cast<unsigned short*>(X);
        for (int e = t; e < 32768; e += blockDim.x) Xh[e] = (unsigned short)(__float_as_uint((float)(e & 255)) >> 16);
    }
    // identity, K-major no-swizzle: element (r,k) at (k/4)*2048 + r*16 + (k%4)*4
    if (c.kind == 0) {
        for (int e = t; e < 128 * 8; e += blockDim.x) {
            const int r = e / 8, k = e % 8;
            *reinterpret_cast<float*>(I + (k >> 2) * 2048 + r * 16 + (k & 3) * 4) = (r == k) ? 1.0f : 0.0f;
        }
    } else {
        for (int e = t; e < 128 * 16; e += blockDim.x) {
            const int r = e / 16, k = e % 16;
            *reinterpret_cast<unsigned short*>(I + (k >> 3) * 2048 + r * 16 + (k & 7) * 2) = (r == k) ? 0x3f80 : 0;
        }
    }

    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *slot;
    {   // sentinel in the accumulator: a dropped MMA leaves 777 behind
        uint32_t v = __float_as_uint(777.0f);
        for (int cc = 0; cc < 32; ++cc)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tm + ((uint32_t)(warp * 32) << 16) + cc), "r"(v) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (t == 0) {
        const uint64_t dx = make_desc(smem_u32(X), c.lbo, c.sbo, c.layout);
        const uint64_t di = make_desc(smem_u32(I), 2048, 128, 0);
        if (c.kind == 0) {
            if (c.which == 0) mma(tm, di, dx, c.idesc, 0);   // D[m][n] = sum_k I[m][k] B[n][k] = B(n, k = m) for m < 8
            else mma(tm, dx, di, c.idesc, 0);                // D[m][n] = sum_k A[m][k] I[n][k] = A(m, k = n) for n < 8
        } else {
            if (c.which == 0) mma_f16(tm, di, dx, c.idesc, 0);
            else mma_f16(tm, dx, di, c.idesc, 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tm + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int cc = 0; cc < 32; ++cc) D[t * 32 + cc] = __uint_as_float(r[cc]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tm));
}

static uint32_t idesc(int M, int N, int amn, int bmn, int fmt = 2) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

int main() {
    float* dD;
    cudaMalloc(&dD, 128 * 32 * 4);
    const int smem = 1024 + 65536 + 4096;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    struct T { const char* name; Cfg c; } tests[] = {
        {"A MN BASE32B lbo=4096 sbo=512", {idesc(128, 32, 1, 0), 1, 4096, 512, 1, 0}},
        {"A MN BASE32B lbo=512 sbo=4096", {idesc(128, 32, 1, 0), 1, 512, 4096, 1, 0}},
        {"B MN BASE32B lbo=4096 sbo=512", {idesc(128, 32, 0, 1), 0, 4096, 512, 1, 0}},
        {"A MN SW64 lbo=2048 sbo=512", {idesc(128, 32, 1, 0), 1, 2048, 512, 4, 0}},
        {"A MN SW32 lbo=2048 sbo=256", {idesc(128, 32, 1, 0), 1, 2048, 256, 6, 0}},
        {"A MN SW128 lbo=8192 sbo=1024", {idesc(128, 32, 1, 0), 1, 8192, 1024, 2, 0}},
    };
    float* h = new float[128 * 32];
    for (auto& t : tests) {
        cudaMemset(dD, 0xff, 128 * 32 * 4);
        probe<<<1, 128, smem>>>(t.c, dD);
        cudaError_t e = cudaDeviceSynchronize();
        printf("== %s: %s\n", t.name, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(h, dD, 128 * 32 * 4, cudaMemcpyDeviceToHost);
        for (int k = 0; k < 8; ++k) {
            printf("  k=%d:", k);
            if (t.c.which == 0) for (int n = 0; n < 32; ++n) printf(" %g", h[k * 32 + n]);
            else for (int m = 0; m < 128; ++m) printf(" %g", h[m * 32 + k]);
            printf("\n");
        }
    }
    return 0;
}
