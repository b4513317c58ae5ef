// Fused training kernel, tensor-core variant: every contraction of the forward AND the backward pass runs on tcgen05
// (5th-generation tensor cores, accumulators in TMEM) in 3xTF32 split precision, which keeps fp32-level accuracy (the
// 1e-5 gradient gate) while taking all multiply-accumulates off the CUDA cores.  What is left for the SM's issue slots
// is the per-sample work: gather, Fourier features, SnakeAlt and its derivative, the hi/lo splits and the scatter --
// about a quarter of the instructions of the FFMA2 kernel (sample_backward_v2.cu).
//
//   tile = 128 samples = the 128 TMEM lanes; TPS threads share one sample (thread (s, q) owns 32/TPS of the 32 hidden
//   columns), so a CTA is 128*TPS threads.  The kernel owns all 512 TMEM columns of its SM (one CTA per SM).
//
//   Activations never touch shared memory on the forward path: the A operand of every sample-major contraction is
//   read by the tensor core straight from TMEM (tcgen05.mma with [a_tmem]; an SS-mode MMA of this shape moves 5 KB of
//   shared memory per instruction and is paced by that, not by the math).  Each layer input lives in TMEM as its tf32
//   "hi" part and the exact remainder "lo" (hi + lo == the fp32 value), written with tcgen05.st by the thread that
//   produced it, so the pair is operand and stash at once:
//       [  0, 32)  working accumulator: z_l in the forward, dh_l / d(features) in the backward
//       [ 32,224)  h_1, h_2, h_3 as (hi | lo), 64 columns each; the layer-0 input (K0p <= 56 columns) borrows the
//                  h_2 (hi) and h_3 (lo) slots, which are free until layers 2 / 3 are reached; in the backward the h_3
//                  slot carries dz_l (hi | lo) (h_3 is read out for its weight-gradient operand first)
//       [224,480)  dW_l^T accumulators, one 64-column block per layer ([A dz_hi | A dz_lo]), kept across the whole
//                  persistent loop
//   The fp32 copy of the layer-0 input (re-split for dW_0 in the backward) lives in shared memory (written in the input
//   stage, read after the forward pass: block barriers in between).
//   Forward, layer l:      z_l = h_l W_l^T          A = h_l (TMEM),            B = W_l   K-major panels (smem)
//   Backward, layer l:     dh_l = dz_l W_l          A = dz_l (TMEM),           B = W_l^T K-major panels (smem)
//                          dW_l^T += [h_l | 1]^T dz_l   A = h_l MN-major (smem), B = dz_l MN-major (smem), K = samples
//   S'(z_l) stays in registers between the forward and the backward (32/TPS per layer and thread).
//   The weight-gradient MMAs are committed to their own mbarrier and only waited for when their operand buffers are
//   about to be overwritten, so they run behind the dz epilogue of the layer and the input stage of the next tile.
//   The ones column appended to h_l (row 63 of the accumulator) makes the bias gradient part of the same MMA.
//
//   Weight-gradient operands: the A block is [h hi g0 | h hi g1 | h lo g0 | h lo g1] (four 32-column groups), read
//   as ONE M = 128 operand, and the B block is [dz hi | dz lo] read as ONE N = 64 operand, so a single MMA per K step
//   yields hi*dz_hi (rows 0..63, columns 0..31), hi*dz_lo (rows 0..63, columns 32..63), lo*dz_hi and the negligible
//   lo*dz_lo (rows 64..127); the flush adds the four quadrants.  (Round 1 issued two N = 32 MMAs per K step, reading the
//   4 KB A block twice: the weight-gradient MMAs are shared-memory-bandwidth bound, 6 KB instead of 10 KB per K step.)
//
//   tf32 operands read MN-major (contraction over the samples) only work in the SWIZZLE_128B_BASE32B layout on sm_100a
//   (measured, profiles/microbench/umma_addr.cu: every other layout type returns zeros for kind::tf32 with a transposed operand):
//       element (mn, k) at (mn/32)*LBO + (k/4)*SBO + (k%4)*128 + (((mn%32)/8) ^ (k%4))*32 + (mn%8)*4
//   i.e. a row-major [sample][32 columns] block whose 32-byte chunks are XOR-swizzled by (sample & 3).
//
// Layer-0 columns are permuted as in the forward kernel: k = [features (Cp) | xyz | Fourier | zero pad to K0p].
#include "sample_backward.cuh"
#include "tc_panels.cuh"

#include <stdlib.h>

#ifdef LFGC_PHASE_TIMING
__device__ unsigned long long g_btc_cycles[16];
__device__ unsigned long long g_btc_time[4];   // globaltimer: [0] entry (CTA 0), [1] latest exit of the last launch; [2], [3]: the launch before
__device__ __forceinline__ unsigned long long btc_gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define BT_DECL if (threadIdx.x == 0 && blockIdx.x == 0) { g_btc_time[2] = g_btc_time[0]; g_btc_time[3] = g_btc_time[1]; g_btc_time[0] = btc_gtimer(); g_btc_time[1] = 0; } long long _tl = clock64(); unsigned long long _tp[16] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0};
#define BT_MARK(i) { long long _n = clock64(); _tp[i] += (unsigned long long)(_n - _tl); _tl = _n; }
#define BT_FLUSH() { if (threadIdx.x == 0) { for (int _i = 0; _i < 16; ++_i) atomicAdd(&g_btc_cycles[_i], _tp[_i]); atomicMax(&g_btc_time[1], btc_gtimer()); } }
#else
#define BT_DECL
#define BT_MARK(i)
#define BT_FLUSH()
#endif

namespace lfgc {
namespace btc {

constexpr int HP = 32;
constexpr int TILE = 128;
constexpr int LMAX = 4;
constexpr int kPanelW = HP * 16;     // bytes per K chunk of a forward weight operand (32 rows)
constexpr int kBlkMN = TILE * 128;   // bytes of one MN-major [128 samples][32 columns] block
constexpr int kOnesRow = 63;         // accumulator row that receives the bias gradient (column 31 of MN group 1)

// TMEM column map
constexpr int cAcc = 0, cA1 = 32, cDz = 160, cW = 224, kWAcc = 64, kTmemCols = 512;   // cDz aliases the h_3 slot
__host__ __device__ constexpr int colA_hi(int l) { return l == 0 ? cA1 + 64 : cA1 + 64 * (l - 1); }
__host__ __device__ constexpr int colA_lo(int l) { return l == 0 ? cA1 + 128 : cA1 + 64 * (l - 1) + 32; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE descriptor (cute::UMMA::SmemDescriptor bit layout, version 1)
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
// MN-major, SWIZZLE_128B_BASE32B: LBO = bytes between 32-column groups, SBO = bytes between 4-sample atoms
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((kBlkMN >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((512 >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// kind::tf32, fp32 accumulate, M = 128
__host__ __device__ constexpr uint32_t idesc(int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t id, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(id), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);  // round to nearest tf32
    lo = x - hi;                                                         // exact in fp32
}
__device__ __forceinline__ void split4(const float4& a, float4& hi, float4& lo) {
    split_tf32(a.x, hi.x, lo.x);
    split_tf32(a.y, hi.y, lo.y);
    split_tf32(a.z, hi.z, lo.z);
    split_tf32(a.w, hi.w, lo.w);
}

// Bounded wait: a lost completion traps (reported as a launch failure) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
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
        if (clock64() - t0 > 2000000000ll) __trap();   // ~1 s
    }
}

// ---- TMEM <-> registers (32 lanes x 32 bit, N consecutive columns) ----------------------------------------------------
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ld<4>(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
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
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

template <int N>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_st<4>(uint32_t taddr, const float (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3]))
                 : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<8>(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<16>(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::
            "r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// A operand from TMEM (lane = sample, 8 consecutive 32-bit columns per K = 8 step), B from shared memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t id, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(id), "r"(accumulate)
        : "memory");
}
// One lane of a converged warp (elect.sync).  Issuing tcgen05.mma under `if (threadIdx.x == 0)` makes the compiler wrap
// every MMA in an ELECT / BRA.U.ANY serialisation loop (~60 cycles per instruction, measured); under a warp-uniform
// branch + elect.sync it emits the bare UTCMMA sequence.
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- shared-memory layout (bytes) ---------------------------------------------------------------------------------------
struct Layout {
    int ctrl, bias, wf, yx, Hm, Dm, X0, Wf, Wb, total;
    int Np0;       // rows of the layer-0 backward weight operand (feature columns, multiple of 16)
    int wb0;       // bytes of the layer-0 backward weight operand (one of hi / lo)
    int wfBytes, wbBytes;  // bytes of one of hi / lo
};

__host__ __device__ inline Layout make_layout(const SampleParams& P, int K0p, int tps) {
    Layout o;
    o.Np0 = (P.Cp + 15) & ~15;
    o.wb0 = (HP / 4) * o.Np0 * 16;
    o.wfBytes = (K0p + (P.L - 1) * HP) / 4 * kPanelW;
    o.wbBytes = o.wb0 + (P.L - 1) * (HP / 4) * kPanelW;
    int p = 0;
    o.ctrl = p; p += 64;                      // mbarriers: [0,8) forward / dh, [8,16) dW; [24,28) TMEM base
    o.bias = p; p += LMAX * HP * 4;
    o.wf = p;   p += (HP + 4) * 4;
    o.yx = p;   p += tps * TILE * 4;          // partial outputs of the threads sharing a sample
    p = (p + 1023) & ~1023;
    o.Hm = p;   p += 4 * kBlkMN;              // MN-major [h hi g0 | h hi g1 (+ones) | h lo g0 | h lo g1]
    o.Dm = p;   p += 2 * kBlkMN;              // MN-major dz_l (hi | lo); parameter staging at start-up, flush scratch
    o.X0 = p;   p += (K0p / 4) * TILE * 16;   // fp32 layer-0 input, [K chunk][sample] float4 (re-split for dW_0)
    o.Wf = p;   p += 2 * o.wfBytes;
    o.Wb = p;   p += 2 * o.wbBytes;
    o.total = p;
    return o;
}

// byte offset of column j (0..31) of sample s inside an MN-major block
__device__ __forceinline__ int mn_off(int s, int j) { return s * 128 + ((((j >> 3) ^ s) & 3) << 5) + ((j & 7) << 2); }

template <int FUSED, int TPS>
__global__ void __launch_bounds__(TILE * TPS, 1) backward_tc_kernel(const __grid_constant__ BwdArgs A, const int K0p) {
    LFGC_PDL_PROLOGUE();
    constexpr int CW = HP / TPS;   // hidden columns per thread
    constexpr int NT = TILE * TPS;
    extern __shared__ __align__(1024) unsigned char smem[];
    BT_DECL
    const SampleParams& P = A.P;
    const Layout Lo = make_layout(P, K0p, TPS);
    float* bias = reinterpret_cast<float*>(smem + Lo.bias);
    float* wfs = reinterpret_cast<float*>(smem + Lo.wf);
    float* yx = reinterpret_cast<float*>(smem + Lo.yx);
    unsigned char* Hm = smem + Lo.Hm;
    unsigned char* DmHi = smem + Lo.Dm;
    unsigned char* DmLo = DmHi + kBlkMN;
    float4* X0s = reinterpret_cast<float4*>(smem + Lo.X0);
    unsigned char* WfHi = smem + Lo.Wf;
    unsigned char* WfLo = WfHi + Lo.wfBytes;
    unsigned char* WbHi = smem + Lo.Wb;
    unsigned char* WbLo = WbHi + Lo.wbBytes;
    const uint32_t barA = smem_u32(smem + Lo.ctrl);        // forward layers and dh
    const uint32_t barB = barA + 8;                        // weight-gradient MMAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Lo.ctrl + 24);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp >> 2;                  // which share of the columns
    const int s = ((warp & 3) << 5) | lane;   // sample within the tile = TMEM lane
    const int col0 = q * CW;                  // first hidden column of this thread
    const int H = P.H, in0 = P.in0, L = P.L;
    const int nfix = 3 + 6 * P.F;
    const int Cp = P.Cp;

    // ---- sample position and layer-0 gathers -------------------------------------------------------------------------------
    uint64_t sample_base = A.sample_offset;
    if (FUSED && A.step_dev) sample_base += (uint64_t)(*A.step_dev) * A.step_stride;
    const bool small_volume = A.n_voxels <= 0xFFFFFFFFull;   // 32-bit index arithmetic (every shipped volume)
    auto load_position = [&](int64_t sg, bool valid, float& cx, float& cy, float& cz, float& aux) {
        cx = cy = cz = aux = 0.0f;
        if (!valid) return;
        if (FUSED && !A.coords) {
            const unsigned long long v = A.explicit_idx ? (unsigned long long)A.explicit_idx[sg]
                                                        : philox_voxel(A.seed, sample_base + (uint64_t)sg, A.n_voxels);
            int i, j, k;
            if (small_volume) {
                const unsigned v32 = (unsigned)v, r2 = (unsigned)A.R[2], r12 = (unsigned)A.R[1] * r2;
                const unsigned ii = v32 / r12, rem = v32 - ii * r12, jj = rem / r2;
                i = (int)ii;
                j = (int)jj;
                k = (int)(rem - jj * r2);
            } else {
                const unsigned long long r12 = (unsigned long long)A.R[1] * A.R[2];
                i = (int)(v / r12);
                j = (int)((v / A.R[2]) % A.R[1]);
                k = (int)(v % A.R[2]);
            }
            cx = normalized_coord((float)i, A.max_idx[0], A.scales[0]);
            cy = normalized_coord((float)j, A.max_idx[1], A.scales[1]);
            cz = normalized_coord((float)k, A.max_idx[2], A.scales[2]);
            aux = __ldg(A.volume + v);
        } else {  // caller-supplied positions; aux = target value (fused) or d(loss)/d(out)
            cx = __ldg(A.coords + 3 * sg);
            cy = __ldg(A.coords + 3 * sg + 1);
            cz = __ldg(A.coords + 3 * sg + 2);
            aux = __ldg(A.grad_out + sg);
        }
    };
    float cx, cy, cz, aux;
    Corners Kc;
    float4 gv[8], gu[8];   // two adjacent feature chunks (8 channels) of this thread, gathered at the 8 corners
    // Feature chunks are taken in adjacent pairs (2p, 2p + 1), p = q, q + TPS, ...: 32 contiguous bytes per corner, fetched
    // with ONE 256-bit load (LDG.E.256, sm_100+) where the rows allow it -- half the load instructions and L1 wavefronts of
    // the divergent gather, which is what bounds the input stage.
    const bool wide_rows = (Cp & 7) == 0 && (reinterpret_cast<uintptr_t>(A.grid) & 31) == 0;
    auto gather_pair = [&](int p) {
        if (wide_rows) {
#pragma unroll
            for (int cc = 0; cc < 8; ++cc)
                asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=f"(gv[cc].x), "=f"(gv[cc].y), "=f"(gv[cc].z), "=f"(gv[cc].w), "=f"(gu[cc].x), "=f"(gu[cc].y),
                               "=f"(gu[cc].z), "=f"(gu[cc].w)
                             : "l"(A.grid + Kc.off[cc] + 8 * p));
        } else {
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) gv[cc] = ldg_f4(A.grid + Kc.off[cc] + 8 * p);
            if (4 * (2 * p + 1) < Cp) {
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) gu[cc] = ldg_f4(A.grid + Kc.off[cc] + 8 * p + 4);
            }
        }
    };
    auto issue_gather = [&]() { gather_pair(q); };
    // ---- one-time setup ------------------------------------------------------------------------------------------------------
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(barA), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(barB), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    const bool have_image = A.panel_image != nullptr;
    {
        // zero what must read as zero / finite: the second column group of the MN-major activation block (only the ones
        // column and the layer-0 columns >= 32 are ever written there) and, when they are built here, the weight panels
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = threadIdx.x; e < kBlkMN / 16; e += NT) {
            reinterpret_cast<float4*>(Hm + kBlkMN)[e] = z4;
            reinterpret_cast<float4*>(Hm + 3 * kBlkMN)[e] = z4;
        }
        if (have_image) {
            // ready-made operands (tc_panels.cuh, kept current by the optimiser step): two straight 128-bit copies, all
            // loads of a thread in flight together
            const float4* hsrc = reinterpret_cast<const float4*>(A.panel_image);
            float4* hdst = reinterpret_cast<float4*>(smem + Lo.bias);
            for (int e = threadIdx.x; e < kTcHdrCopy / 16; e += NT) hdst[e] = __ldg(hsrc + e);
            // the panels are first needed by the MMAs after the first tile's input stage: asynchronous copies (LDGSTS),
            // waited for just before that stage's closing barrier
            const unsigned char* psrc = A.panel_image + kTcHdrBytes;
            const uint32_t pdst = smem_u32(smem + Lo.Wf);
            const int n16 = (Lo.total - Lo.Wf) / 16;
            for (int e = threadIdx.x; e < n16; e += NT)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pdst + 16u * (uint32_t)e), "l"(psrc + 16 * (size_t)e) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            for (int e = threadIdx.x; e < (Lo.total - Lo.Wf) / 16; e += NT) reinterpret_cast<float4*>(smem + Lo.Wf)[e] = z4;
            // the packed parameter block, staged through shared memory with independent loads (one L2 round trip)
            float* stage = reinterpret_cast<float*>(DmHi);
            if ((reinterpret_cast<uintptr_t>(A.mlp) & 15) == 0) {
                const int n4 = A.pcount >> 2;
                for (int e4 = threadIdx.x; e4 < n4; e4 += NT)
                    reinterpret_cast<float4*>(stage)[e4] = __ldg(reinterpret_cast<const float4*>(A.mlp) + e4);
                if ((int)threadIdx.x < (A.pcount & 3)) stage[4 * n4 + threadIdx.x] = __ldg(A.mlp + 4 * n4 + threadIdx.x);
            } else {
                for (int e = threadIdx.x; e < A.pcount; e += NT) stage[e] = __ldg(A.mlp + e);
            }
        }
    }
    // The first tile's loads go out before the rest of the setup (the Philox draw, the volume read and the 16 gathers are
    // three dependent memory round trips) but AFTER its light part above, which covers the latency of the step-counter load
    // the Philox counter depends on.
    if ((int64_t)blockIdx.x * TILE < A.n) {
        const int64_t sg0 = (int64_t)blockIdx.x * TILE + s;
        load_position(sg0, sg0 < A.n, cx, cy, cz, aux);
        make_corners(P, cx, cy, cz, Kc);
        issue_gather();
    }

    __syncthreads();
    BT_MARK(12)  // setup: zero fill, parameter staging, TMEM allocation
    // ones column (column 31 of group 1 of the hi block): bias gradient row of every dW accumulator
    for (int r = threadIdx.x; r < TILE; r += NT) *reinterpret_cast<float*>(Hm + kBlkMN + mn_off(r, 31)) = 1.0f;
    if (!have_image) {
        const float* stage = reinterpret_cast<const float*>(DmHi);
        for (int l = 0; l < L; ++l) {
            const int K = l == 0 ? in0 : H;
            const float* W = stage + mlp_w_off(l, in0, H);
            const int fbase = (l == 0 ? 0 : (K0p + (l - 1) * HP) / 4) * kPanelW;
            const int bbase = l == 0 ? 0 : Lo.wb0 + (l - 1) * (HP / 4) * kPanelW;
            const int brows = l == 0 ? Lo.Np0 : HP;
            // warp = output row j (stride NT/32), lane = input column r; everything that depends on the column only is
            // hoisted, so an element costs a load, the split and four stores with constant strides
            constexpr int JS = NT / 32;   // a multiple of 4: (j & 3) is the same for every j of a warp
            for (int r = lane; r < K; r += 32) {
                int k = r;
                if (l == 0) k = r < nfix ? Cp + r : r - nfix;  // permuted layer-0 columns
                const bool back = l > 0 || k < Cp;             // backward operand: only the feature columns of layer 0
                const float* src = W + warp * K + r;
                unsigned char* fh = WfHi + fbase + (k >> 2) * kPanelW + (k & 3) * 4 + warp * 16;            // forward: B[n = j][K = k]
                unsigned char* bh = WbHi + bbase + k * 16 + (warp & 3) * 4 + (warp >> 2) * brows * 16;      // backward: B[n = k][K = j]
                const int fstep = JS * 16, bstep = (JS >> 2) * brows * 16;
                const ptrdiff_t flo = WfLo - WfHi, blo = WbLo - WbHi;
                for (int j = warp; j < H; j += JS) {
                    float hi, lo;
                    split_tf32(*src, hi, lo);
                    *reinterpret_cast<float*>(fh) = hi;
                    *reinterpret_cast<float*>(fh + flo) = lo;
                    if (back) {
                        *reinterpret_cast<float*>(bh) = hi;
                        *reinterpret_cast<float*>(bh + blo) = lo;
                    }
                    src += JS * K;
                    fh += fstep;
                    bh += bstep;
                }
            }
            const float* b = stage + mlp_b_off(l, in0, H);
            for (int j = threadIdx.x; j < HP; j += NT) bias[l * HP + j] = j < H ? b[j] : 0.0f;
        }
        const float* wf = stage + mlp_wf_off(L, in0, H);
        for (int j = threadIdx.x; j < HP; j += NT) wfs[j] = j < H ? wf[j] : 0.0f;
        if (threadIdx.x == 0) wfs[HP] = wf[H];
    }
    BT_MARK(13)  // setup: weight panels
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's lane quarter
    const float bf = wfs[HP];
    const uint32_t aHm = smem_u32(Hm), aDmHi = smem_u32(DmHi), aDmLo = smem_u32(DmLo);
    const uint32_t aWfHi = smem_u32(WfHi), aWfLo = smem_u32(WfLo), aWbHi = smem_u32(WbHi), aWbLo = smem_u32(WbLo);
    uint32_t parA = 0, parB = 0;
    bool pendingB = false;     // weight-gradient MMAs in flight: their operand buffers must not be overwritten yet
    bool first_tile = true;

    float accWf[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) accWf[i] = 0.0f;
    float accbf = 0.0f, loss_part = 0.0f;

    // this thread's 4 columns [j0, j0+4) of an MN-major block pair
    auto put_mn = [&](unsigned char* hi_base, unsigned char* lo_base, int j0, const float4& hi, const float4& lo) {
        const int off = mn_off(s, j0);
        *reinterpret_cast<float4*>(hi_base + off) = hi;
        *reinterpret_cast<float4*>(lo_base + off) = lo;
    };
    BT_MARK(0)  // setup

    const int64_t ntiles = (A.n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- input stage: the TPS threads of a sample take the K chunks c = q, q + TPS, ... ----------------------------
        const int64_t sg = tile * TILE + s;
        const bool valid = sg < A.n;
        if (!first_tile) {   // the first tile's position and gathers were issued ahead of the one-time setup
            load_position(sg, valid, cx, cy, cz, aux);
            make_corners(P, cx, cy, cz, Kc);
            issue_gather();
        }
        // one K chunk (4 columns) of the layer-0 input: tf32 hi / lo operand columns and the fp32 copy
        auto emit = [&](int c, const float (&v4)[4]) {
            float hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) split_tf32(v4[i], hi[i], lo[i]);
            tmem_st<4>(trow + colA_hi(0) + 4 * c, hi);
            tmem_st<4>(trow + colA_lo(0) + 4 * c, lo);
            X0s[c * TILE + s] = make_float4(v4[0], v4[1], v4[2], v4[3]);
        };
        // positional columns first: their sincos work runs while the gathers are in flight.  A chunk's four columns take
        // two packed evaluations (sincos_fast2: two arguments each, sin or cos picked per column)
        for (int c = (Cp >> 2) + q; c < K0p / 4; c += TPS) {
            float arg[4], v4[4];
            int kind[4];   // 0: the argument itself (xyz), 1: sin, 2: cos, 3: zero pad
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = 4 * c + i - Cp;   // 0..2 xyz, then per frequency [sin x y z | cos x y z]
                arg[i] = 0.0f;
                kind[i] = 3;
                if (r < 3) {
                    arg[i] = r == 0 ? cx : (r == 1 ? cy : cz);
                    kind[i] = 0;
                } else if (r < nfix) {
                    const int f = (r - 3) / 6, m = (r - 3) - 6 * f;
                    const int ax = m >= 3 ? m - 3 : m;
                    const float coord = ax == 0 ? cx : (ax == 1 ? cy : cz);
                    arg[i] = __fmul_rn(coord, P.omega[f]);   // argument rounded to fp32 first
                    kind[i] = m >= 3 ? 2 : 1;
                }
            }
            float sn[4], cs[4];
            sincos_fast2(arg[0], arg[1], sn[0], cs[0], sn[1], cs[1]);
            sincos_fast2(arg[2], arg[3], sn[2], cs[2], sn[3], cs[3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) v4[i] = kind[i] == 0 ? arg[i] : (kind[i] == 1 ? sn[i] : (kind[i] == 2 ? cs[i] : 0.0f));
            emit(c, v4);
        }
        // feature chunks in adjacent pairs; the first pair was issued above
        for (int p = q; 8 * p < Cp; p += TPS) {
            const int c = 2 * p, c2 = c + 1;
            const bool two = 4 * c2 < Cp;
            if (p != q) gather_pair(p);
            float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                a[0] = fmaf(gv[cc].x, Kc.w[cc], a[0]);
                a[1] = fmaf(gv[cc].y, Kc.w[cc], a[1]);
                a[2] = fmaf(gv[cc].z, Kc.w[cc], a[2]);
                a[3] = fmaf(gv[cc].w, Kc.w[cc], a[3]);
            }
            emit(c, a);
            if (two) {
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    b[0] = fmaf(gu[cc].x, Kc.w[cc], b[0]);
                    b[1] = fmaf(gu[cc].y, Kc.w[cc], b[1]);
                    b[2] = fmaf(gu[cc].z, Kc.w[cc], b[2]);
                    b[3] = fmaf(gu[cc].w, Kc.w[cc], b[3]);
                }
                emit(c2, b);
            }
        }
        tmem_st_wait();
        if (first_tile && have_image) {   // the weight panels have landed; make them visible to the tensor core's reads
            asm volatile("cp.async.wait_all;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        BT_MARK(1)  // input stage
        __syncthreads();
        BT_MARK(2)  // barriers

        // ---- forward ------------------------------------------------------------------------------------------------------------
        float hs[CW], gs[LMAX - 1][CW], glast[CW];   // S'(z_l): l < L-1 in gs[l], the last layer's in glast
#pragma unroll
        for (int l = 0; l < LMAX; ++l) {
            if (l < L) {
                if (warp == 0 && elect_one()) {   // the MMAs of this layer, issued by one lane (see elect_one)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const int nk = (l == 0 ? K0p : HP) / 8;
                    const uint32_t boff = (uint32_t)((l == 0 ? 0 : (K0p + (l - 1) * HP) / 4) * kPanelW);
                    constexpr uint32_t id = idesc(HP, 0, 0);
                    const uint32_t ah = tmem + colA_hi(l), al = tmem + colA_lo(l);
                    const uint64_t bh0 = desc_k(aWfHi + boff, kPanelW, 128), bl0 = desc_k(aWfLo + boff, kPanelW, 128);
                    for (int ks = 0; ks < nk; ++ks) {
                        const uint64_t adv = (uint64_t)((ks * 2 * kPanelW) >> 4);
                        mma_tf32_ts(tmem + cAcc, ah + 8 * ks, bh0 + adv, id, ks > 0 ? 1u : 0u);
                        mma_tf32_ts(tmem + cAcc, al + 8 * ks, bh0 + adv, id, 1u);
                        mma_tf32_ts(tmem + cAcc, ah + 8 * ks, bl0 + adv, id, 1u);
                    }
                    mma_commit(barA);
                }
                BT_MARK(3)  // MMA issue
                mbar_wait(barA, parA);
                parA ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                BT_MARK(4)  // MMA wait

                float z[CW];
                tmem_ld<CW>(trow + cAcc + col0, z);
                const float* bl_ = bias + l * HP + col0;
#pragma unroll
                for (int i = 0; i < CW; i += 2) {
                    float g0, g1;
                    snake_and_grad_fast2(z[i] + bl_[i], z[i + 1] + bl_[i + 1], hs[i], hs[i + 1], g0, g1);
                    if (l + 1 < L) {
                        if (l + 1 < LMAX) { gs[l < LMAX - 1 ? l : 0][i] = g0; gs[l < LMAX - 1 ? l : 0][i + 1] = g1; }
                    } else {
                        glast[i] = g0;
                        glast[i + 1] = g1;
                    }
                }
                if (l + 1 < L) {
                    float hi[CW], lo[CW];
#pragma unroll
                    for (int i = 0; i < CW; ++i) split_tf32(hs[i], hi[i], lo[i]);
                    tmem_st<CW>(trow + colA_hi(l + 1) + col0, hi);
                    tmem_st<CW>(trow + colA_lo(l + 1) + col0, lo);
                    tmem_st_wait();
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    BT_MARK(5)  // forward epilogue
                    __syncthreads();
                    BT_MARK(2)
                }
            }
        }

        // ---- output, loss, dz_{L-1} --------------------------------------------------------------------------------------------
        {
            float yp = 0.0f;
#pragma unroll
            for (int i = 0; i < CW; ++i) yp = fmaf(hs[i], wfs[col0 + i], yp);
            yx[q * TILE + s] = yp;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // also: every thread has read its z_{L-1}
        float dy;
        {
            float y = bf;
#pragma unroll
            for (int t = 0; t < TPS; ++t) y += yx[t * TILE + s];
            if (FUSED) {
                const float e = y - aux;
                dy = valid ? A.loss_scale2 * e : 0.0f;
                if (q == 0 && valid) loss_part = fmaf(e, e, loss_part);
                if (A.log_sigma && valid) {   // Gaussian likelihood with per-sample log sigma (VariationalDropoutLoss,
                    const float w = expf(-2.0f * __ldg(A.log_sigma + sg));   // Variational_Dropout_Layer.py:24-30)
                    dy *= w;
                    if (A.dlog_sigma && q == 0) A.dlog_sigma[sg] = A.loss_scale2 * (1.0f - e * e * w);
                }
            } else {
                dy = valid ? aux : 0.0f;
            }
        }
#pragma unroll
        for (int i = 0; i < CW; ++i) accWf[i] = fmaf(dy, hs[i], accWf[i]);
        if (q == 0) accbf += dy;
        float dz[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i) dz[i] = dy * wfs[col0 + i] * glast[i];
        BT_MARK(6)  // output + loss

        // ---- backward ----------------------------------------------------------------------------------------------------------
#pragma unroll
        for (int l = LMAX - 1; l >= 0; --l) {
            if (l < L) {
                if (pendingB) {   // the previous weight-gradient MMAs still read the operand buffers written below
                    mbar_wait(barB, parB);
                    parB ^= 1u;
                    pendingB = false;
                }
                BT_MARK(11)  // dW wait
                // h_l -> MN-major operand (group 0; the layer-0 input also fills its columns >= 32 of group 1).  Read BEFORE
                // dz_l is written: dz shares the TMEM slot of h_3 (same lanes and columns per thread)
                if (l > 0) {
                    float hi[CW], lo[CW];
                    tmem_ld<CW>(trow + colA_hi(l) + col0, hi);
                    tmem_ld<CW>(trow + colA_lo(l) + col0, lo);
#pragma unroll
                    for (int c = 0; c < CW / 4; ++c)
                        put_mn(Hm, Hm + 2 * kBlkMN, col0 + 4 * c, make_float4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]),
                               make_float4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]));
                } else {
                    for (int c = q; c < K0p / 4; c += TPS) {
                        const float4 hv = X0s[c * TILE + s];
                        float4 hi, lo;
                        split4(hv, hi, lo);
                        const int g = (4 * c) >> 5;   // 32-column group
                        put_mn(Hm + g * kBlkMN, Hm + (2 + g) * kBlkMN, (4 * c) & 31, hi, lo);
                    }
                }
                // dz_l: TMEM operand of dh_l (hi | lo) and MN-major operand of dW_l
                {
                    float hi[CW], lo[CW];
#pragma unroll
                    for (int i = 0; i < CW; ++i) split_tf32(dz[i], hi[i], lo[i]);
                    tmem_st<CW>(trow + cDz + col0, hi);
                    tmem_st<CW>(trow + cDz + HP + col0, lo);
#pragma unroll
                    for (int c = 0; c < CW / 4; ++c)
                        put_mn(DmHi, DmLo, col0 + 4 * c, make_float4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]),
                               make_float4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]));
                }
                tmem_st_wait();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                BT_MARK(7)  // backward operand staging
                __syncthreads();
                BT_MARK(2)
                if (warp == 0 && elect_one()) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // dh_l = dz_l W_l  (layer 0: only the feature columns)
                    {
                        const uint32_t id = l > 0 ? idesc(HP, 0, 0) : idesc(Lo.Np0, 0, 0);
                        const uint32_t rows16 = (uint32_t)(l > 0 ? HP : Lo.Np0) * 16;
                        const uint32_t boff = (uint32_t)(l > 0 ? Lo.wb0 + (l - 1) * (HP / 4) * kPanelW : 0);
                        const uint64_t bh0 = desc_k(aWbHi + boff, rows16, 128), bl0 = desc_k(aWbLo + boff, rows16, 128);
                        for (int ks = 0; ks < HP / 8; ++ks) {
                            const uint64_t adv = (uint64_t)((ks * 2 * rows16) >> 4);
                            mma_tf32_ts(tmem + cAcc, tmem + cDz + 8 * ks, bh0 + adv, id, ks > 0 ? 1u : 0u);
                            mma_tf32_ts(tmem + cAcc, tmem + cDz + HP + 8 * ks, bh0 + adv, id, 1u);
                            mma_tf32_ts(tmem + cAcc, tmem + cDz + 8 * ks, bl0 + adv, id, 1u);
                        }
                        mma_commit(barA);
                    }
                    // dW_l^T += [h_l | 1]^T [dz_hi | dz_lo], contraction over the 128 samples (8 per MMA); rows 0..63 of the
                    // accumulator collect hi * dz, rows 64..127 lo * dz; columns 0..31 dz_hi, 32..63 dz_lo
                    {
                        constexpr uint32_t id = idesc(2 * HP, 1, 1);
                        const uint32_t d = tmem + cW + l * kWAcc;
                        const uint64_t a0 = desc_mn(aHm), b0 = desc_mn(aDmHi);   // DmLo follows DmHi: the second N group
                        for (int ks = 0; ks < TILE / 8; ++ks) {
                            const uint64_t adv = (uint64_t)((ks * 1024) >> 4);
                            mma_tf32(d, a0 + adv, b0 + adv, id, (first_tile && ks == 0) ? 0u : 1u);
                        }
                        mma_commit(barB);
                    }
                }
                pendingB = true;
                BT_MARK(3)
                mbar_wait(barA, parA);
                parA ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                BT_MARK(4)
                if (l > 0) {
                    float dh[CW];
                    tmem_ld<CW>(trow + cAcc + col0, dh);
#pragma unroll
                    for (int i = 0; i < CW; ++i) dz[i] = dh[i] * gs[l > 0 ? l - 1 : 0][i];
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    BT_MARK(8)  // dz
                } else {
                    // scatter d(features) into the grid gradient
                    for (int c = q; 4 * c < Cp; c += TPS) {
                        float d[4];
                        tmem_ld<4>(trow + cAcc + 4 * c, d);   // warp-collective: outside the per-sample predicate
                        if (valid) {
#pragma unroll
                            for (int cc = 0; cc < 8; ++cc) {
                                const float w = Kc.w[cc];
                                if (w != 0.0f)
                                    red_add_v4(A.grad_grid + Kc.off[cc] + 4 * c, make_float4(d[0] * w, d[1] * w, d[2] * w, d[3] * w));
                            }
                        }
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    BT_MARK(9)  // scatter
                }
            }
        }
        first_tile = false;
    }

    // ---- flush: weight-gradient accumulators, final-layer gradient, loss ------------------------------------------------------
    if (pendingB) {
        mbar_wait(barB, parB);
        parB ^= 1u;
        pendingB = false;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // destination: this CTA's workspace slice (reduced later, fixed order), or -- lfgc_train_step_accumulate -- the running
    // sums themselves, added to with atomics (148 CTAs x 3.4 k reductions at the very end of the kernel cost less than the
    // reduction launch they replace in the data-parallel step)
    float* const aout = A.atomic_out;
    float* dst = aout ? aout + (size_t)(blockIdx.x % (unsigned)A.atomic_slices) * A.pstride : A.partial + (size_t)blockIdx.x * A.pstride;
    auto put = [&](int idx, float v) {
        if (aout) atomicAdd(dst + idx, v);
        else dst[idx] = v;
    };
    // Every accumulator row (TMEM lane) passes through shared memory once: S[layer][row 0..127][33 floats].  With a row
    // stride of 33 both the lane-per-row stores and the row-fastest reads are bank-conflict free, and the global stores
    // become one compact loop over the packed parameter order in which all warps take part (the former register-to-global
    // flush was ~850 instructions of straight-line code run by half of the warps: 8.6 k cycles, mostly instruction fetch).
    constexpr int kSRow = 33;
    float* S = reinterpret_cast<float*>(Hm);                 // Hm and Dm are contiguous: 96 KB
    float* red = S + LMAX * TILE * kSRow;                    // [128][36]: 32 Wf partials | bf | loss
    for (int l = q; l < L; l += TPS) {
        float r[32], r2[32];
        tmem_ld<32>(trow + cW + l * kWAcc, r);               // columns: dz hi
        tmem_ld<32>(trow + cW + l * kWAcc + HP, r2);         //          dz lo
        float* row = S + (l * TILE + s) * kSRow;
#pragma unroll
        for (int j = 0; j < 32; ++j) row[j] = r[j] + r2[j];
    }
#pragma unroll
    for (int i = 0; i < CW; ++i) red[s * 36 + col0 + i] = accWf[i];
    if (q == 0) {
        red[s * 36 + 32] = accbf;
        red[s * 36 + 33] = FUSED ? loss_part : 0.0f;
    }
    __syncthreads();
    BT_MARK(14)  // flush: accumulators to shared memory
    for (int l = 0; l < L; ++l) {
        const int Kin = l == 0 ? in0 : H;
        const int woff = mlp_w_off(l, in0, H), boff = mlp_b_off(l, in0, H);
        const float* Sl = S + l * TILE * kSRow;
        // warp = output row j, lane = input column k: coalesced stores, conflict-free reads, pointer increments only
        for (int k = lane; k < Kin; k += 32) {
            const int row = l > 0 ? k : (k < nfix ? Cp + k : k - nfix);   // layer 0: permuted columns
            const float* src = Sl + row * kSRow + warp;                  // rows 0..63 hold hi * dz, rows 64..127 lo * dz
            int o = woff + warp * Kin + k;
            for (int j = warp; j < H; j += NT / 32) {
                put(o, src[0] + src[64 * kSRow]);
                src += NT / 32;
                o += (NT / 32) * Kin;
            }
        }
        if (warp == (l & 3) && lane < H)   // bias: the ones row
            put(boff + lane, Sl[kOnesRow * kSRow + lane] + Sl[(kOnesRow + 64) * kSRow + lane]);
    }
    BT_MARK(15)  // flush: dW to the partial slice
    {
        // column sums over the 128 samples: 7 row groups x 36 columns in parallel, then 7 partials per column
        float* red2 = red + TILE * 36;
        {
            const int col = threadIdx.x % 36, rg = threadIdx.x / 36;
            if (rg < 7 && col < 34) {
                float t = 0.0f;
                for (int r = rg; r < TILE; r += 7) t += red[r * 36 + col];
                red2[rg * 36 + col] = t;
            }
        }
        __syncthreads();
        if (threadIdx.x < 34) {
            float t = 0.0f;
#pragma unroll
            for (int g = 0; g < 7; ++g) t += red2[g * 36 + threadIdx.x];
            const int wfo = mlp_wf_off(L, in0, H);
            if (threadIdx.x < 32) {
                if ((int)threadIdx.x < H) put(wfo + threadIdx.x, t);
            } else if (threadIdx.x == 32) {
                put(wfo + H, t);
            } else {
                put(A.pcount, t);
            }
        }
    }
    if (aout && A.ann_n > 0) {
        // early announcement (lfgc_peer_announce): every thread's reductions are performed device-wide before the CTA takes
        // its ticket; the CTA that draws the last one knows all of this rank's sums are complete and tells the peers
        __threadfence();
        __syncthreads();
        if (warp == 0) {
            int last = 0;
            if (lane == 0) {
                const int ticket = atomicAdd(A.ann_ticket, 1);
                if (ticket == (int)gridDim.x - 1) {
                    *A.ann_ticket = 0;
                    __threadfence();
                    last = 1;
                }
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) announce_epoch(A, lane);
        }
    }
    BT_MARK(10)  // flush
    BT_FLUSH()
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
}

template <int FUSED, int TPS>
static int launch_tps(BwdArgs& A, int K0p, float* grad_mlp, int accumulate, void* workspace, size_t workspace_bytes,
                      cudaStream_t st) {
    const Layout Lo = make_layout(A.P, K0p, TPS);
    if (Lo.total > max_smem_optin()) return 1;
    if ((size_t)A.pcount * sizeof(float) > (size_t)2 * kBlkMN) return 1;   // parameter staging buffer
    auto kern = backward_tc_kernel<FUSED, TPS>;
    LFGC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Lo.total));
    const int64_t ntiles = (A.n + TILE - 1) / TILE;
    int64_t grid = sm_count();   // the kernel owns the SM's TMEM: one CTA per SM
    if (grid > ntiles) grid = ntiles;
    const size_t need = (size_t)grid * A.pstride * sizeof(float);
    if (workspace_bytes < need) return fail(LFGC_E_WORKSPACE, "backward workspace too small: %zu < %zu", workspace_bytes, need);
    A.partial = reinterpret_cast<float*>(workspace);
    (void)launch_pdl(kern, dim3((unsigned)grid), dim3(TILE * TPS), (size_t)(Lo.total), st, A, K0p);
    LFGC_LAUNCH_OK();
    if (A.atomic_out) {   // the kernel added its sums itself
        A.nslices = 0;
        return LFGC_OK;
    }
    finish_partials(A, (int)grid, grad_mlp, accumulate, FUSED ? A.loss_sum : nullptr, st);
    if (!A.defer_reduce) LFGC_LAUNCH_OK();
    return LFGC_OK;
}

}  // namespace btc

// Returns LFGC_OK after launching, 1 when the shape is not covered by the tensor-core kernel, or an error code.
int launch_backward_tc(BwdArgs& A, int fused, float* grad_mlp, int accumulate, void* workspace, size_t workspace_bytes,
                       cudaStream_t st) {
    const SampleParams& P = A.P;
    if (P.H > btc::HP || P.L < 1 || P.L > btc::LMAX) return 1;
    const int K0 = P.Cp + 3 + 6 * P.F;
    const int K0p = (K0 + 7) & ~7;
    if (K0p > 56) return 1;   // two 32-column MN groups, the last column of the second one is the ones column
    int tps = 2;
    if (const char* e = getenv("LFGC_TC_TPS")) { const int v = atoi(e); if (v == 2 || v == 4) tps = v; }  // tuning aid
    if (fused) {
        return tps == 4 ? btc::launch_tps<1, 4>(A, K0p, grad_mlp, accumulate, workspace, workspace_bytes, st)
                        : btc::launch_tps<1, 2>(A, K0p, grad_mlp, accumulate, workspace, workspace_bytes, st);
    }
    return tps == 4 ? btc::launch_tps<0, 4>(A, K0p, grad_mlp, accumulate, workspace, workspace_bytes, st)
                    : btc::launch_tps<0, 2>(A, K0p, grad_mlp, accumulate, workspace, workspace_bytes, st);
}

}  // namespace lfgc

namespace lfgc {
__global__ void tc_panel_build_kernel(const float* __restrict__ mlp, unsigned char* __restrict__ img, const TcPanelMap M,
                                      int pcount) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pcount) tc_panel_store(img, M, i, mlp[i]);
}
}  // namespace lfgc

extern "C" size_t lfgc_tc_panel_bytes(const lfgc_model_desc* m) {
    lfgc::SampleParams P;
    if (!m || lfgc::fill_sample_params(m, 0, P)) return 0;
    return (size_t)lfgc::make_tc_panel_map(P).total_bytes;
}

extern "C" int lfgc_tc_panel_build(const lfgc_model_desc* m, const float* mlp, float* image, void* stream) {
    using namespace lfgc;
    SampleParams P;
    if (!m || !mlp || !image) return fail(LFGC_E_INVALID, "tc_panel_build: null pointer");
    const int rc = fill_sample_params(m, 0, P);
    if (rc) return rc;
    const TcPanelMap M = make_tc_panel_map(P);
    if (M.total_bytes == 0) return fail(LFGC_E_UNSUPPORTED, "tc_panel_build: the tensor-core kernel does not cover this model");
    if (reinterpret_cast<uintptr_t>(image) & 15) return fail(LFGC_E_INVALID, "tc_panel_build: the image must be 16-byte aligned");
    const int pcount = (int)lfgc_mlp_param_count(m);
    LFGC_CUDA_OK(cudaMemsetAsync(image, 0, (size_t)M.total_bytes, (cudaStream_t)stream));
    tc_panel_build_kernel<<<(pcount + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mlp, reinterpret_cast<unsigned char*>(image), M, pcount);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

#ifdef LFGC_PHASE_TIMING
extern "C" int lfgc_btc_timing(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (out16) cudaMemcpyFromSymbol(out16, g_btc_cycles, sizeof(unsigned long long) * 16);
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_btc_cycles, z, sizeof(z));
    }
    return 0;
}
extern "C" int lfgc_btc_time(unsigned long long* out2, int reset) {
    cudaDeviceSynchronize();
    if (out2) cudaMemcpyFromSymbol(out2, g_btc_time, sizeof(unsigned long long) * 4);
    if (reset) {
        unsigned long long z[4] = {0, 0, 0, 0};
        cudaMemcpyToSymbol(g_btc_time, z, sizeof(z));
    }
    return 0;
}
#endif
