// Fused sample kernel (forward), tensor-core variant: the hidden-layer contractions run on tcgen05 (5th-generation
// tensor cores, accumulators in TMEM) in 3xTF32 split precision, which keeps fp32-level accuracy (the 1e-5 parity
// gate) while taking the multiply-accumulates off the CUDA cores; what remains for the SM's issue slots is the
// per-sample epilogue (bias, SnakeAlt, hi/lo split) -- about a third of the instructions of the FFMA2 kernel.
//
//   tile = 128 samples = the 128 TMEM lanes of one accumulator; thread t of the CTA owns sample t.
//   A operand (activations, M=128 x K): shared memory, K-major, no swizzle, stored as 16-byte "chunk panels":
//       element (row r, k) at (k/4)*2048 + r*16 + (k%4)*4 bytes  -> thread r writes its row with conflict-free
//       128-bit stores; LBO (between the two K chunks of one K=8 MMA) = 2048 B, SBO (between 8-row groups) = 128 B.
//   B operand (weights, N=32 x K): same panel layout with 512-byte panels, staged once per CTA.
//   Both operands exist twice, as tf32 "hi" (round-to-nearest to 10 mantissa bits) and "lo" = x - hi, and each
//   K=8 step issues hi*hi + lo*hi + hi*lo (the dropped lo*lo term is 2^-22 relative).
//   Per layer: one thread issues the 12 MMAs and commits to an mbarrier; all threads wait, read their accumulator row
//   with tcgen05.ld (32 columns), apply bias + SnakeAlt, split and write the next A operand.
//   A kernel that allocates TMEM is limited to ONE resident CTA per SM (measured: cudaOccupancy reports 1 for any
//   kernel containing tcgen05.alloc, profiles/microbench/tmem_occ.cu), so latency hiding happens inside the CTA: it holds G
//   independent "tile groups" of 128 threads, each with its own operand buffer, 32 TMEM columns, mbarrier and tile
//   loop, sharing one copy of the weight panels; while one group waits for its MMAs the others run their epilogues.
//
// Layer-0 column order is permuted so that the gathered feature vectors land on whole chunks: k = [features (Cp) |
// xyz (3) | Fourier (6F) | zero pad]; the layer-0 weight panel is built with the same permutation.
#include "sample_common.cuh"

#include <stdlib.h>

#ifdef LFGC_PHASE_TIMING
__device__ unsigned long long g_tc_cycles[8];
#define TC_DECL long long _tl = clock64(); unsigned long long _tp[8] = {0,0,0,0,0,0,0,0};
#define TC_MARK(i) { long long _n = clock64(); _tp[i] += (unsigned long long)(_n - _tl); _tl = _n; }
#define TC_FLUSH() { if (threadIdx.x == 0) for (int _i = 0; _i < 8; ++_i) atomicAdd(&g_tc_cycles[_i], _tp[_i]); }
#else
#define TC_DECL
#define TC_MARK(i)
#define TC_FLUSH()
#endif

namespace lfgc {
namespace tc {

constexpr int HP = 32;
constexpr int TILE = 128;
constexpr int kPanelA = TILE * 16;  // bytes per K chunk of the activation operand
constexpr int kPanelB = HP * 16;    // bytes per K chunk of a weight operand

struct Args {
    SampleParams P;
    const float* coords;
    const float* axis[3];
    int R1, R2;
    int64_t first;
    int64_t n;
    const float* grid;
    const float* mlp;
    float* out;
    int K0p;  // padded layer-0 contraction length (multiple of 8)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
    return d;
}

// instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 32
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HP >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);  // round to nearest tf32
    lo = x - hi;                                                         // exact in fp32
}

// Bounded wait: a lost completion traps (reported as a launch failure) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

// One lane of a converged warp (elect.sync): under `if (t == 0)` the compiler wraps every tcgen05.mma in an ELECT /
// BRA.U.ANY serialisation loop (~60 cycles per instruction); under a warp-uniform branch + elect.sync it does not.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

struct Layout {
    int bias, Wf, Bhi, Blo, Ahi, Alo, total;  // byte offsets
};

__host__ __device__ inline Layout make_layout(const SampleParams& P, int K0p, int G) {
    Layout o;
    int p = 64;  // [0, 8G) one mbarrier per group, [56,60) TMEM base address
    o.bias = p; p += P.L * HP * 4;
    o.Wf = p;   p += (HP + 4) * 4;
    p = (p + 127) & ~127;
    const int kB = (K0p + (P.L - 1) * HP) / 4 * kPanelB;
    o.Bhi = p;  p += kB;
    o.Blo = p;  p += kB;
    const int chunksA = (K0p > HP ? K0p : HP) / 4;
    o.Ahi = p;  p += G * chunksA * kPanelA;   // group g at Ahi + g * chunksA * kPanelA
    o.Alo = p;  p += G * chunksA * kPanelA;
    o.total = p;
    return o;
}

template <int MODE, int G>
__global__ void __launch_bounds__(TILE * G, 1) sample_forward_tc_kernel(const __grid_constant__ Args A) {
    constexpr int NT = TILE * G;                                            // threads per CTA
    constexpr int kTmemCols = G == 1 ? 32 : (G == 2 ? 64 : (G <= 4 ? 128 : 256));  // power of two >= 32 G
    extern __shared__ __align__(128) unsigned char smem[];
    const SampleParams& P = A.P;
    const int K0p = A.K0p;
    const Layout Lo = make_layout(P, K0p, G);
    float* bias = reinterpret_cast<float*>(smem + Lo.bias);
    float* Wf = reinterpret_cast<float*>(smem + Lo.Wf);
    unsigned char* Bhi = smem + Lo.Bhi;
    unsigned char* Blo = smem + Lo.Blo;
    const int g = threadIdx.x / TILE;                                       // tile group of this thread
    const int t = threadIdx.x % TILE;                                       // sample / TMEM lane within the group
    const int chunksA = (K0p > HP ? K0p : HP) / 4;
    unsigned char* Ahi = smem + Lo.Ahi + g * chunksA * kPanelA;
    unsigned char* Alo = smem + Lo.Alo + g * chunksA * kPanelA;
    const uint32_t bar = smem_u32(smem + 8 * g);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 56);
    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(TILE) : "memory"); };

    TC_DECL
    const int warp = t >> 5;                                                // warp within the group = TMEM lane quarter
    const int H = P.H, in0 = P.in0, L = P.L;
    const int nfix = 3 + 6 * P.F;  // xyz + Fourier columns

    // ---- one-time setup: barrier, TMEM, weights (hi/lo panels), biases --------------------------------------------
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    {
        const int kB_floats = (K0p + (L - 1) * HP) * HP;
        float* bh = reinterpret_cast<float*>(Bhi);
        float* bl = reinterpret_cast<float*>(Blo);
        for (int e = threadIdx.x; e < kB_floats; e += NT) { bh[e] = 0.0f; bl[e] = 0.0f; }
    }
    __syncthreads();
    for (int l = 0; l < L; ++l) {
        const int K = l == 0 ? in0 : H;
        const float* W = A.mlp + mlp_w_off(l, in0, H);
        const int base_chunk = l == 0 ? 0 : (K0p + (l - 1) * HP) / 4;
        for (int e = threadIdx.x; e < H * K; e += NT) {
            const int j = e / K, r = e - j * K;
            int k = r;
            if (l == 0) k = r < nfix ? P.Cp + r : r - nfix;  // permuted layer-0 columns
            float hi, lo;
            split_tf32(__ldg(W + e), hi, lo);
            const int off = (base_chunk + (k >> 2)) * kPanelB + j * 16 + (k & 3) * 4;
            *reinterpret_cast<float*>(Bhi + off) = hi;
            *reinterpret_cast<float*>(Blo + off) = lo;
        }
        const float* b = A.mlp + mlp_b_off(l, in0, H);
        for (int j = threadIdx.x; j < HP; j += NT) bias[l * HP + j] = j < H ? __ldg(b + j) : 0.0f;
    }
    {
        const float* wf = A.mlp + mlp_wf_off(L, in0, H);
        for (int j = threadIdx.x; j < HP; j += NT) Wf[j] = j < H ? __ldg(wf + j) : 0.0f;
        if (threadIdx.x == 0) Wf[HP] = __ldg(wf + H);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_alloc = *tmem_slot;
    const uint32_t tmem_base = tmem_alloc + 32u * g;                        // this group's 32 accumulator columns
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float bf = Wf[HP];
    const bool clampo = (P.flags & LFGC_F_CLAMP) != 0;
    const uint32_t aHi = smem_u32(Ahi), aLo = smem_u32(Alo), bHi = smem_u32(Bhi), bLo = smem_u32(Blo);
    uint32_t parity = 0;
    TC_MARK(0)  // setup

    const int64_t ntiles = (A.n + TILE - 1) / TILE;
    for (int64_t tile = (int64_t)blockIdx.x * G + g; tile < ntiles; tile += (int64_t)gridDim.x * G) {
        // ---- input stage: thread = sample; writes the layer-0 operand row (hi and lo) ----------------------------------
        const int64_t s = tile * TILE + t;
        {
            float cx = 0.f, cy = 0.f, cz = 0.f;
            if (s < A.n) {
                if (MODE == 0) {
                    cx = __ldg(A.coords + 3 * s);
                    cy = __ldg(A.coords + 3 * s + 1);
                    cz = __ldg(A.coords + 3 * s + 2);
                } else {
                    const int64_t v = A.first + s;
                    int i, j, k;
                    if (v < 0x100000000ll && (int64_t)A.R2 * A.R1 < 0x100000000ll) {   // 32-bit index arithmetic (every shipped volume): two divisions instead of
                        const unsigned v32 = (unsigned)v, r2 = (unsigned)A.R2, r12 = r2 * (unsigned)A.R1;   // three 64-bit ones
                        const unsigned ii = v32 / r12, rem = v32 - ii * r12, jj = rem / r2;
                        i = (int)ii;
                        j = (int)jj;
                        k = (int)(rem - jj * r2);
                    } else {
                        k = (int)(v % A.R2);
                        j = (int)((v / A.R2) % A.R1);
                        i = (int)(v / ((int64_t)A.R2 * A.R1));
                    }
                    cx = __ldg(A.axis[0] + i);
                    cy = __ldg(A.axis[1] + j);
                    cz = __ldg(A.axis[2] + k);
                }
            }
            Corners K;
            make_corners(P, cx, cy, cz, K);
            const bool wide_rows = (P.Cp & 7) == 0 && (reinterpret_cast<uintptr_t>(A.grid) & 31) == 0;
            // two channel chunks per batch: 16 independent 128-bit loads in flight per thread (the gather's L2 round
            // trips are the largest stall of this kernel; ncu source view)
            auto emit = [&](const float4& a, int chunk) {
                float4 hi, lo;
                split_tf32(a.x, hi.x, lo.x);
                split_tf32(a.y, hi.y, lo.y);
                split_tf32(a.z, hi.z, lo.z);
                split_tf32(a.w, hi.w, lo.w);
                *reinterpret_cast<float4*>(Ahi + chunk * kPanelA + t * 16) = hi;
                *reinterpret_cast<float4*>(Alo + chunk * kPanelA + t * 16) = lo;
            };
            for (int c4 = 0; c4 < P.Cp; c4 += 8) {
                const bool two = c4 + 4 < P.Cp;
                float4 v[8], u[8];
                if (wide_rows) {   // both chunks with one 256-bit load per corner (LDG.E.256, sm_100+)
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                     : "=f"(v[c].x), "=f"(v[c].y), "=f"(v[c].z), "=f"(v[c].w), "=f"(u[c].x), "=f"(u[c].y),
                                       "=f"(u[c].z), "=f"(u[c].w)
                                     : "l"(A.grid + K.off[c] + c4));
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[c] = ldg_f4(A.grid + K.off[c] + c4);
                    if (two) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) u[c] = ldg_f4(A.grid + K.off[c] + c4 + 4);
                    }
                }
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    a.x = fmaf(v[c].x, K.w[c], a.x);
                    a.y = fmaf(v[c].y, K.w[c], a.y);
                    a.z = fmaf(v[c].z, K.w[c], a.z);
                    a.w = fmaf(v[c].w, K.w[c], a.w);
                }
                emit(a, c4 >> 2);
                if (two) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        b.x = fmaf(u[c].x, K.w[c], b.x);
                        b.y = fmaf(u[c].y, K.w[c], b.y);
                        b.z = fmaf(u[c].z, K.w[c], b.z);
                        b.w = fmaf(u[c].w, K.w[c], b.w);
                    }
                    emit(b, (c4 >> 2) + 1);
                }
            }
            auto put = [&](int k, float x) {
                float hi, lo;
                split_tf32(x, hi, lo);
                const int off = (k >> 2) * kPanelA + t * 16 + (k & 3) * 4;
                *reinterpret_cast<float*>(Ahi + off) = hi;
                *reinterpret_cast<float*>(Alo + off) = lo;
            };
            int k = P.Cp;
            put(k++, cx);
            put(k++, cy);
            put(k++, cz);
            // Fourier features two arguments at a time (sincos_fast2); frequencies in pairs so that no lane is wasted
            for (int f = 0; f < P.F; f += 2) {
                const float om = P.omega[f];
                const bool pair = f + 1 < P.F;
                const float om2 = pair ? P.omega[f + 1] : om;
                float sx, cxx, sy, cyy, sz, czz, sx2, cx2, sy2, cy2, sz2, cz2;
                sincos_fast2(__fmul_rn(cx, om), __fmul_rn(cy, om), sx, cxx, sy, cyy);   // arguments rounded to fp32 first
                sincos_fast2(__fmul_rn(cz, om), __fmul_rn(cx, om2), sz, czz, sx2, cx2);
                put(k++, sx); put(k++, sy); put(k++, sz);
                put(k++, cxx); put(k++, cyy); put(k++, czz);
                if (pair) {
                    sincos_fast2(__fmul_rn(cy, om2), __fmul_rn(cz, om2), sy2, cy2, sz2, cz2);
                    put(k++, sx2); put(k++, sy2); put(k++, sz2);
                    put(k++, cx2); put(k++, cy2); put(k++, cz2);
                }
            }
            for (; k < K0p; ++k) put(k, 0.0f);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        TC_MARK(1)  // input stage
        group_sync();
        TC_MARK(2)  // barrier

        float y = 0.0f;
        for (int l = 0; l < L; ++l) {
            // ---- one thread issues the 3xTF32 MMAs of this layer --------------------------------------------------------
            if (warp == 0 && elect_one()) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int nk = (l == 0 ? K0p : HP) / 8;
                const uint32_t boff = (uint32_t)((l == 0 ? 0 : (K0p + (l - 1) * HP) / 4) * kPanelB);
                for (int ks = 0; ks < nk; ++ks) {
                    const uint64_t ah = make_desc(aHi + ks * 2 * kPanelA, kPanelA, 128);
                    const uint64_t al = make_desc(aLo + ks * 2 * kPanelA, kPanelA, 128);
                    const uint64_t bh = make_desc(bHi + boff + ks * 2 * kPanelB, kPanelB, 128);
                    const uint64_t bl = make_desc(bLo + boff + ks * 2 * kPanelB, kPanelB, 128);
                    mma_tf32(tmem_base, ah, bh, ks > 0 ? 1u : 0u);
                    mma_tf32(tmem_base, al, bh, 1u);
                    mma_tf32(tmem_base, ah, bl, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                             : "memory");
            }
            TC_MARK(3)  // MMA issue
            mbar_wait(bar, parity);
            TC_MARK(4)  // MMA wait
            parity ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

            // ---- epilogue: this thread's accumulator row ----------------------------------------------------------------
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32"
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
                " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(tmem_row));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const float* bl_ = bias + l * HP;
            // SnakeAlt value with the single-polynomial, packed-pair evaluation (snake_value_fast2): this epilogue is what
            // bounds the kernel (issue-active 75 %, r1 ncu), so its instruction count is the reconstruction rate
            if (l + 1 < L) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bl_ + 4 * c);
                    float h0, h1, h2, h3;
                    snake_value_fast2(__uint_as_float(r[4 * c + 0]) + b4.x, __uint_as_float(r[4 * c + 1]) + b4.y, h0, h1);
                    snake_value_fast2(__uint_as_float(r[4 * c + 2]) + b4.z, __uint_as_float(r[4 * c + 3]) + b4.w, h2, h3);
                    float4 hi, lo;
                    split_tf32(h0, hi.x, lo.x);
                    split_tf32(h1, hi.y, lo.y);
                    split_tf32(h2, hi.z, lo.z);
                    split_tf32(h3, hi.w, lo.w);
                    *reinterpret_cast<float4*>(Ahi + c * kPanelA + t * 16) = hi;
                    *reinterpret_cast<float4*>(Alo + c * kPanelA + t * 16) = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bl_ + 4 * c);
                    const float4 w4 = *reinterpret_cast<const float4*>(Wf + 4 * c);
                    float h0, h1, h2, h3;
                    snake_value_fast2(__uint_as_float(r[4 * c + 0]) + b4.x, __uint_as_float(r[4 * c + 1]) + b4.y, h0, h1);
                    snake_value_fast2(__uint_as_float(r[4 * c + 2]) + b4.z, __uint_as_float(r[4 * c + 3]) + b4.w, h2, h3);
                    y = fmaf(h0, w4.x, y);
                    y = fmaf(h1, w4.y, y);
                    y = fmaf(h2, w4.z, y);
                    y = fmaf(h3, w4.w, y);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            TC_MARK(5)  // epilogue
            group_sync();  // operand rows of every thread are written / every accumulator row has been read
            TC_MARK(2)
        }
        if (s < A.n) {
            float v = y + bf;
            if (clampo) v = fminf(fmaxf(v, -1.0f), 1.0f);
            A.out[s] = v;
        }
    }

    TC_FLUSH()
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_alloc), "n"(kTmemCols));
}

template <int MODE, int G>
static int launch_g(const Args& A, cudaStream_t st) {
    const Layout Lo = make_layout(A.P, A.K0p, G);
    auto kern = sample_forward_tc_kernel<MODE, G>;
    LFGC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Lo.total));
    const int64_t ntiles = (A.n + TILE - 1) / TILE;
    int64_t grid = sm_count();                       // TMEM kernels: one CTA per SM
    const int64_t need = (ntiles + G - 1) / G;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, TILE * G, Lo.total, st>>>(A);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

template <int MODE>
static int launch(const Args& A, cudaStream_t st) {
    int G = 5;
    if (const char* e = getenv("LFGC_TC_GROUPS")) { const int v = atoi(e); if (v >= 1 && v <= 5) G = v; }  // tuning aid
    const int64_t ntiles = (A.n + TILE - 1) / TILE;
    while (G > 1 && (make_layout(A.P, A.K0p, G).total > max_smem_optin() || (int64_t)sm_count() * (G - 1) >= ntiles)) --G;
    if (make_layout(A.P, A.K0p, G).total > max_smem_optin()) return 1;
    switch (G) {
        case 5: return launch_g<MODE, 5>(A, st);
        case 4: return launch_g<MODE, 4>(A, st);
        case 3: return launch_g<MODE, 3>(A, st);
        case 2: return launch_g<MODE, 2>(A, st);
        default: return launch_g<MODE, 1>(A, st);
    }
}

}  // namespace tc

// Returns LFGC_OK after launching, 1 when the shape is not covered by the tensor-core kernel, or an error code.
int launch_forward_tc(const SampleParams& P, const float* coords, const float* const axis[3], int R1, int R2,
                      int64_t first, int64_t n, const float* grid, const float* mlp, float* out, cudaStream_t st) {
    if (P.H > tc::HP || P.L < 1) return 1;
    const int K0 = P.Cp + 3 + 6 * P.F;
    const int K0p = (K0 + 7) & ~7;
    if (K0p > 128) return 1;
    tc::Args A;
    A.P = P;
    A.coords = coords;
    A.axis[0] = axis ? axis[0] : nullptr;
    A.axis[1] = axis ? axis[1] : nullptr;
    A.axis[2] = axis ? axis[2] : nullptr;
    A.R1 = R1;
    A.R2 = R2;
    A.first = first;
    A.n = n;
    A.grid = grid;
    A.mlp = mlp;
    A.out = out;
    A.K0p = K0p;
    return coords ? tc::launch<0>(A, st) : tc::launch<1>(A, st);
}

}  // namespace lfgc

#ifdef LFGC_PHASE_TIMING
extern "C" int lfgc_tc_timing(unsigned long long* out8, int reset) {
    cudaDeviceSynchronize();
    if (out8) cudaMemcpyFromSymbol(out8, g_tc_cycles, sizeof(unsigned long long) * 8);
    if (reset) {
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(g_tc_cycles, z, sizeof(z));
    }
    return 0;
}
#endif
