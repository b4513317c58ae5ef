// "Grid step" kernel: everything of one optimiser step that is NOT per-sample, for mask-free models, in ONE launch:
//
//     [sum of several gradient buffers, e.g. of data-parallel ranks mapped into this address space]
//     reduction of the fused training kernel's MLP-gradient partial sums (+ the loss)
//     synthesis adjoint, finest -> coarsest level        (autograd of Feature_Grid_Model.decode_volume,
//                                                         model/Feature_Grid_Model.py:102-108, training/training.py:137)
//     [SmallifyLoss weight term  2 w sum coeff^2 -> g]    (model/Smallify_Dropout.py:29,37)
//     Adam over coefficients and MLP                      (torch.optim.Adam, training/training.py:199,232)
//     synthesis, coarsest -> finest, of the UPDATED coefficients: the channels-last grid the next step gathers from,
//     and the cleared gradient accumulator it scatters into
//
// It replaces six dependent launches (reduce_partials, 2 x idwt_level_bwd, adam, 2 x idwt_level at the shipped 2-level
// configurations: 21.6 us + 4.3 us of a 75.8 us step on B200, round-1 launch list) whose cost is launch / ramp latency,
// not work: the whole working set is 0.45 MB.  Round 1 tried the same fusion as a cooperative 296-CTA launch with
// grid-wide barriers and measured it SLOWER than the separate launches (81.9 vs 75.8 us per step; removed).  This kernel
// needs no grid-wide barrier at all because the wavelet transform never mixes channels
// (wavelet_transform/Torch_Wavelet_Transform.py:100-104, groups = C):
//
//     CTA c < C owns channel c: its slice of the grid gradient, of every coefficient tensor, of m / v and of the decoded
//     grid.  The whole per-channel pyramid lives in shared memory (76 KB at C16/G15) and the phases are separated by
//     __syncthreads.  Each synthesis / adjoint level is evaluated SEPARABLY (x, y, z passes of n_taps MACs per output
//     instead of the direct (n_taps/2)^3 x 8 sum: 6.4x fewer MACs for db2), which is what makes one CTA per channel enough.
//     CTAs >= C: the MLP block (partial-sum reduction in fixed order = deterministic, Adam).
//
// Summation order differs from wavelet.cu's direct sums (separable evaluation); both are checked against the fp64 oracle.
// Models whose per-channel pyramid does not fit in shared memory (G > ~22) or with live mask layers keep the separate
// kernels (FastTrainer falls back; lfgc_grid_step_smem_bytes() == 0 says so).
#include "wavelet_lines.cuh"
#include "tc_panels.cuh"

#include <stdlib.h>

namespace lfgc {
namespace gstep {

constexpr int kThreads = 1024;
constexpr int kMaxCoeff = 8;   // coefficient tensors this kernel handles (pyramids that fit in shared memory have few)


struct LevelPlan {
    int d[3], t[3], off[3];
    int cg_off;                 // offset (floats) of this level's detail bands inside the shared coefficient block
    int m_lo[3], n_m[3];        // synthesis: first pair index and number of pairs along each axis
    FastDiv by_t2, by_d0, by_d0d1;
    PassPlan sx, sy, sz, az, ay, ax;   // synthesis x / y / z, adjoint z^T / y^T / x^T
};

struct Args {
    int n_coeff, C, Cp, ntaps;
    float lo[LFGC_MAX_TAPS], hi[LFGC_MAX_TAPS];
    int d0[3];                        // extent of coefficient tensor 0
    LevelPlan lv[kMaxCoeff];          // [l] for synthesis level l >= 1
    int sL, sX, sY, sC;               // shared-memory partition sizes (floats): low-pass, x-pass, y-pass, coefficients
    int n_srcs;
    const float* grad_grid[LFGC_MAX_PEERS];
    const float* mlp_partials[LFGC_MAX_PEERS];
    int nslices;
    int pstride, pcount;
    float* zero_grid;
    float* grid_cl;
    float* p;
    float* g;
    float* m;
    float* v;
    long long coeff_off[kMaxCoeff];
    long long mlp_off;
    float* loss_out;
    const float* lr;
    int* step;
    AdamCoef c;
    float w2x2;                       // 2 * weight_l2
    int n_mlp_ctas;
    unsigned char* panel_image;       // tensor-core weight operands kept current with the MLP parameters (tc_panels.cuh), nullable
    TcPanelMap pm;
};

// Shared-memory (or, in the host test, heap) partitions of one channel
struct Bufs {
    float* L;    // current low-pass volume / grid gradient / coefficient tensor 0 of this channel
    float* X;    // x-pass intermediate: [4 (a,b)][d0][d1][t2]
    float* Y;    // y-pass intermediate: [2 (a)][d0][t1][t2]
    float* Cg;   // detail bands of every level: gradient after the adjoint, updated value after Adam
};

// Every pass is organised by COLUMNS: a work unit is one 1-D line of the pass (fixed position in the two untouched
// dimensions) times a range of positions along the filtered dimension.  The index decomposition happens once per unit
// (multiply-high instead of integer division), the filter length is a template parameter (taps in registers, loops
// unrolled), and values loaded once are shared: a synthesis position m produces the output pair (2m, 2m+1) - off from the
// same NT/2 inputs of each band; an adjoint position i produces BOTH bands from the same NT outputs.  (A first version
// decomposed a flat index per output element with six integer divisions and a runtime tap loop: 44 us, slower than the
// six launches it replaces; the second, column-organised one still executed ~100 instructions per output: 45 us.)

// band k = 4a + 2b + c, (a, b, c) = filter along (dim0, dim1, dim2), 0 = low / 1 = high; k = 0 is the running low-pass
template <int NT, bool CG = false>
struct Pass {
    // synthesis, x pass: column = (ab, iz, iy), 4 d0 d1 of them
    static __host__ __device__ __forceinline__ void sx(const Args& A, const Bufs& S, const LevelPlan& P, int col, int r0,
                                                       int r1, const float (&flo)[NT], const float (&fhi)[NT]) {
        const int d0d1 = P.d[0] * P.d[1], d2 = P.d[2], dvol = d0d1 * d2;
        const int ab = fdiv(col, P.by_d0d1);
        const int zy = col - ab * d0d1;
        const float* b0 = (ab == 0 ? S.L : S.Cg + P.cg_off + (2 * ab - 1) * dvol) + zy * d2;
        const float* b1 = S.Cg + P.cg_off + (2 * ab) * dvol + zy * d2;
        synth_line<NT, CG>(b0, b1, 1, d2, S.X + col * P.t[2], 1, P.t[2], P.off[2], P.m_lo[2] + r0, P.m_lo[2] + r1, flo, fhi);
    }
    // synthesis, y pass: column = (a, iz, ox), 2 d0 t2 of them
    static __host__ __device__ __forceinline__ void sy(const Args& A, const Bufs& S, const LevelPlan& P, int col, int r0,
                                                       int r1, const float (&flo)[NT], const float (&fhi)[NT]) {
        const int d0 = P.d[0], d1 = P.d[1], t1 = P.t[1], t2 = P.t[2];
        const int az = fdiv(col, P.by_t2);
        const int ox = col - az * t2;
        const int a_ = fdiv(az, P.by_d0);
        const int iz = az - a_ * d0;
        const int plane = d0 * d1 * t2;
        const float* x0 = S.X + (2 * a_) * plane + iz * d1 * t2 + ox;
        synth_line<NT, CG>(x0, x0 + plane, t2, d1, S.Y + az * t1 * t2 + ox, t2, t1, P.off[1], P.m_lo[1] + r0, P.m_lo[1] + r1,
                       flo, fhi);
    }
    // synthesis, z pass: column = (oy, ox), t1 t2 of them
    static __host__ __device__ __forceinline__ void sz(const Args& A, const Bufs& S, const LevelPlan& P, int col, int r0,
                                                       int r1, const float (&flo)[NT], const float (&fhi)[NT]) {
        const int pl = P.t[1] * P.t[2];
        const float* y0 = S.Y + col;
        synth_line<NT, CG>(y0, y0 + P.d[0] * pl, pl, P.d[0], S.L + col, pl, P.t[0], P.off[0], P.m_lo[0] + r0, P.m_lo[0] + r1,
                       flo, fhi);
    }
    // adjoint, z^T pass: column = (oy, ox)
    static __host__ __device__ __forceinline__ void az(const Args& A, const Bufs& S, const LevelPlan& P, int col, int r0,
                                                       int r1, const float (&flo)[NT], const float (&fhi)[NT]) {
        const int pl = P.t[1] * P.t[2];
        adj_line<NT, CG>(S.L + col, pl, P.t[0], P.off[0], S.Y + col, S.Y + P.d[0] * pl + col, pl, r0, r1, flo, fhi);
    }
    // adjoint, y^T pass: column = (a, iz, ox)
    static __host__ __device__ __forceinline__ void ay(const Args& A, const Bufs& S, const LevelPlan& P, int col, int r0,
                                                       int r1, const float (&flo)[NT], const float (&fhi)[NT]) {
        const int d0 = P.d[0], d1 = P.d[1], t1 = P.t[1], t2 = P.t[2];
        const int az_ = fdiv(col, P.by_t2);
        const int ox = col - az_ * t2;
        const int a_ = fdiv(az_, P.by_d0);
        const int iz = az_ - a_ * d0;
        const int plane = d0 * d1 * t2;
        float* o0 = S.X + (2 * a_) * plane + iz * d1 * t2 + ox;
        adj_line<NT, CG>(S.Y + az_ * t1 * t2 + ox, t2, t1, P.off[1], o0, o0 + plane, t2, r0, r1, flo, fhi);
    }
    // adjoint, x^T pass: column = (ab, iz, iy)
    static __host__ __device__ __forceinline__ void ax(const Args& A, const Bufs& S, const LevelPlan& P, int col, int r0,
                                                       int r1, const float (&flo)[NT], const float (&fhi)[NT]) {
        const int d0d1 = P.d[0] * P.d[1], d2 = P.d[2], dvol = d0d1 * d2;
        const int ab = fdiv(col, P.by_d0d1);
        const int zy = col - ab * d0d1;
        float* o0 = (ab == 0 ? S.L : S.Cg + P.cg_off + (2 * ab - 1) * dvol) + zy * d2;   // band (a, b, 0)
        float* o1 = S.Cg + P.cg_off + (2 * ab) * dvol + zy * d2;                          // band (a, b, 1)
        adj_line<NT, CG>(S.X + col * P.t[2], 1, P.t[2], P.off[2], o0, o1, 1, r0, r1, flo, fhi);
    }
};

__host__ __device__ __forceinline__ int coeff_elems(const Args& A, int l) {
    if (l == 0) return A.d0[0] * A.d0[1] * A.d0[2];
    return 7 * A.lv[l].d[0] * A.lv[l].d[1] * A.lv[l].d[2];
}

// run one pass: unit u of `nworkers` strided workers -> (group, column) -> position range
#define LFGC_RUN_PASS(fn, plan_, first_, stride_)                                              \
    {                                                                                          \
        const PassPlan& pp = (plan_);                                                          \
        for (int u = (first_); u < pp.ncols * pp.groups; u += (stride_)) {                     \
            const int grp = fdiv(u, pp.by_ncols);                                              \
            const int col = u - grp * pp.ncols;                                                \
            const int r0_ = grp * pp.per;                                                      \
            const int r1_ = r0_ + pp.per < pp.n ? r0_ + pp.per : pp.n;                         \
            PassT::fn(A, S, P, col, r0_, r1_, flo, fhi);                                       \
        }                                                                                      \
    }

#ifdef __CUDACC__
template <int NT>
__global__ void __launch_bounds__(kThreads, 1) grid_step_kernel(const __grid_constant__ Args A) {
    LFGC_PDL_PROLOGUE();
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_step_size, s_bc2_sqrt;
    __shared__ int s_step;
    const int tid = threadIdx.x;
    if (tid == 0) {
        const int step = *reinterpret_cast<volatile int*>(A.step) + 1;
        float step_size, bc2_sqrt;
        adam_step_scalars(A.c, step, *A.lr, step_size, bc2_sqrt);
        s_step = step;
        s_step_size = step_size;
        s_bc2_sqrt = bc2_sqrt;
    }
    const int c = blockIdx.x;
    using PassT = Pass<NT, false>;
    if (c < A.C) {
        Bufs S;
        S.L = smem;
        S.X = S.L + A.sL;
        S.Y = S.X + A.sX;
        S.Cg = S.Y + A.sY;
        float flo[NT], fhi[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            flo[i] = A.lo[i];
            fhi[i] = A.hi[i];
        }
        const int last = A.n_coeff - 1;
        // ---- this channel's slice of the grid gradient (summed over the ranks in rank order), accumulator cleared ----------
        const int nvox = last >= 1 ? A.lv[last].t[0] * A.lv[last].t[1] * A.lv[last].t[2] : A.d0[0] * A.d0[1] * A.d0[2];
        for (int i = tid; i < nvox; i += kThreads) {
            const long long a = (long long)i * A.Cp + c;
            float gsum = __ldcv(A.grad_grid[0] + a);          // rewritten every step (also by peers): never from L1
            for (int r = 1; r < A.n_srcs; ++r) gsum += __ldcv(A.grad_grid[r] + a);
            S.L[i] = gsum;
            if (A.zero_grid) A.zero_grid[a] = 0.0f;
        }
        if (A.zero_grid && c == A.C - 1) {                      // pad channels of the accumulator stay zero anyway; keep them so
            const int np = A.Cp - A.C;
            for (int i = tid; i < nvox * np; i += kThreads) A.zero_grid[(long long)(i / np) * A.Cp + A.C + i % np] = 0.0f;
        }
        __syncthreads();
        // ---- adjoint, finest -> coarsest --------------------------------------------------------------------------------------
        for (int l = last; l >= 1; --l) {
            const LevelPlan& P = A.lv[l];
            LFGC_RUN_PASS(az, P.az, tid, kThreads)
            __syncthreads();
            LFGC_RUN_PASS(ay, P.ay, tid, kThreads)
            __syncthreads();
            LFGC_RUN_PASS(ax, P.ax, tid, kThreads)
            __syncthreads();
        }
        // ---- Adam on this channel's coefficients ---------------------------------------------------------------------------------
        const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
        for (int l = 0; l <= last; ++l) {
            const int n_l = coeff_elems(A, l);
            float* slots = l == 0 ? S.L : S.Cg + A.lv[l].cg_off;
            const long long base = A.coeff_off[l] + (long long)c * n_l;
            for (int e0 = 0; e0 < n_l; e0 += 4 * kThreads) {       // four elements per thread: 12 loads in flight
                float pi[4], mi[4], vi[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int e = e0 + k * kThreads + tid;
                    if (e < n_l) {
                        pi[k] = A.p[base + e];
                        mi[k] = A.m[base + e];
                        vi[k] = A.v[base + e];
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int e = e0 + k * kThreads + tid;
                    if (e < n_l) {
                        const float gi = fmaf(A.w2x2, pi[k], slots[e]);
                        adam_update(pi[k], gi, mi[k], vi[k], A.c, step_size, bc2_sqrt);
                        A.p[base + e] = pi[k];
                        A.m[base + e] = mi[k];
                        A.v[base + e] = vi[k];
                        A.g[base + e] = gi;
                        slots[e] = pi[k];
                    }
                }
            }
        }
        __syncthreads();
        // ---- synthesis of the updated coefficients, coarsest -> finest ------------------------------------------------------------
        for (int l = 1; l <= last; ++l) {
            const LevelPlan& P = A.lv[l];
            LFGC_RUN_PASS(sx, P.sx, tid, kThreads)
            __syncthreads();
            LFGC_RUN_PASS(sy, P.sy, tid, kThreads)
            __syncthreads();
            LFGC_RUN_PASS(sz, P.sz, tid, kThreads)
            __syncthreads();
        }
        for (int i = tid; i < nvox; i += kThreads) A.grid_cl[(long long)i * A.Cp + c] = S.L[i];
        if (c == A.C - 1) {
            const int np = A.Cp - A.C;
            for (int i = tid; i < nvox * np; i += kThreads) A.grid_cl[(long long)(i / np) * A.Cp + A.C + i % np] = 0.0f;
        }
    } else {
        // ---- MLP block: fixed-order reduction of the partial sums (deterministic), then Adam.  64 parameters x 16 slice
        // groups per CTA: every thread has all its loads in flight at once, the 16 partial sums meet in shared memory -----
        constexpr int PX = 64, SY = kThreads / PX;
        const int px = tid % PX, sy = tid / PX;
        const int j = (c - A.C) * PX + px;
        float acc = 0.0f;
        if (j <= A.pcount) {
            for (int r = 0; r < A.n_srcs; ++r) {
                const float* src = A.mlp_partials[r] + j;
                for (int b = sy; b < A.nslices; b += 4 * SY) {
                    float t0 = __ldcv(src + (size_t)b * A.pstride), t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;
                    if (b + SY < A.nslices) t1 = __ldcv(src + (size_t)(b + SY) * A.pstride);
                    if (b + 2 * SY < A.nslices) t2 = __ldcv(src + (size_t)(b + 2 * SY) * A.pstride);
                    if (b + 3 * SY < A.nslices) t3 = __ldcv(src + (size_t)(b + 3 * SY) * A.pstride);
                    acc += (t0 + t1) + (t2 + t3);
                }
            }
        }
        smem[sy * PX + px] = acc;
        __syncthreads();
        if (sy == 0 && j <= A.pcount) {
            float t = 0.0f;
#pragma unroll
            for (int g = 0; g < SY; ++g) t += smem[g * PX + px];
            if (j < A.pcount) {
                const long long i = A.mlp_off + j;
                float pi = A.p[i], mi = A.m[i], vi = A.v[i];
                adam_update(pi, t, mi, vi, A.c, s_step_size, s_bc2_sqrt);
                A.p[i] = pi;
                A.m[i] = mi;
                A.v[i] = vi;
                A.g[i] = t;
                if (A.panel_image) tc_panel_store(A.panel_image, A.pm, j, pi);   // hi / lo parts at their operand positions
            } else if (A.loss_out) {
                A.loss_out[0] = t;
            }
        }
    }
    // ---- publish the new step count once every CTA has read the old one (ticket scheme of adam_kernel) ---------------------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const int ticket = atomicAdd(A.step + 1, 1);
        if (ticket == (int)gridDim.x - 1) {
            A.step[1] = 0;
            __threadfence();
            A.step[0] = s_step;
        }
    }
}

#endif

}  // namespace gstep
}  // namespace lfgc

using namespace lfgc;

// Level plans and shared-memory partition sizes; returns the total in floats
static size_t gstep_layout(gstep::Args& A, const lfgc_wavelet_desc* w, int nworkers = gstep::kThreads) {
    using namespace gstep;
    A.n_coeff = w->n_coeff;
    A.C = w->C;
    A.ntaps = w->n_taps;
    for (int i = 0; i < LFGC_MAX_TAPS; ++i) {
        A.lo[i] = i < w->n_taps ? w->rec_lo[i] : 0.0f;
        A.hi[i] = i < w->n_taps ? w->rec_hi[i] : 0.0f;
    }
    for (int a = 0; a < 3; ++a) A.d0[a] = w->dims[0][a];
    size_t sL = (size_t)A.d0[0] * A.d0[1] * A.d0[2], sX = 0, sY = 0, sC = 0;
    for (int l = 0; l < kMaxCoeff; ++l) {
        LevelPlan& P = A.lv[l];
        P = LevelPlan();
        if (l < 1 || l >= w->n_coeff) continue;
        for (int a = 0; a < 3; ++a) {
            P.d[a] = w->dims[l][a];
            P.t[a] = w->target[l][a];
            P.off[a] = (2 * P.d[a] + A.ntaps - 2 - P.t[a]) / 2;
            P.m_lo[a] = P.off[a] >> 1;
            P.n_m[a] = ((P.off[a] + P.t[a] - 1) >> 1) - P.m_lo[a] + 1;
        }
        const int d0 = P.d[0], d1 = P.d[1], d2 = P.d[2], t0 = P.t[0], t1 = P.t[1], t2 = P.t[2];
        const size_t dv = (size_t)d0 * d1 * d2, tv = (size_t)t0 * t1 * t2;
        if (dv > sL) sL = dv;
        if (tv > sL) sL = tv;
        const size_t x = 4 * (size_t)d0 * d1 * t2, y = 2 * (size_t)d0 * t1 * t2;
        if (x > sX) sX = x;
        if (y > sY) sY = y;
        P.cg_off = (int)sC;
        sC += 7 * dv;
        P.by_t2 = make_fastdiv((unsigned)t2);
        P.by_d0 = make_fastdiv((unsigned)d0);
        P.by_d0d1 = make_fastdiv((unsigned)(d0 * d1));
        P.sx = make_plan(4 * d0 * d1, P.n_m[2], nworkers);
        P.sy = make_plan(2 * d0 * t2, P.n_m[1], nworkers);
        P.sz = make_plan(t1 * t2, P.n_m[0], nworkers);
        P.az = make_plan(t1 * t2, d0, nworkers);
        P.ay = make_plan(2 * d0 * t2, d1, nworkers);
        P.ax = make_plan(4 * d0 * d1, d2, nworkers);
    }
    auto up4 = [](size_t v) { return (v + 3) & ~(size_t)3; };
    A.sL = (int)up4(sL);
    A.sX = (int)up4(sX);
    A.sY = (int)up4(sY);
    A.sC = (int)up4(sC);
    return (size_t)A.sL + A.sX + A.sY + A.sC;
}

static int gstep_check_desc(const lfgc_wavelet_desc* w) {
    if (!w) return fail(LFGC_E_INVALID, "grid_step: null descriptor");
    if (w->n_coeff < 1 || w->n_coeff > LFGC_MAX_LEVELS || w->C < 1) return fail(LFGC_E_INVALID, "grid_step: bad descriptor");
    if (w->n_coeff > gstep::kMaxCoeff) return fail(LFGC_E_UNSUPPORTED, "grid_step: %d coefficient tensors (<= %d)", w->n_coeff, gstep::kMaxCoeff);
    if (w->n_coeff > 1 && w->n_taps != 2 && w->n_taps != 4)   // what the oracle pins (haar, db2); longer filters: separate kernels
        return fail(LFGC_E_UNSUPPORTED, "grid_step: filter length %d (2 or 4 taps)", w->n_taps);
    return LFGC_OK;
}

extern "C" size_t lfgc_grid_step_smem_bytes(const lfgc_wavelet_desc* w) {
    if (gstep_check_desc(w)) return 0;
    gstep::Args A;
    size_t bytes = gstep_layout(A, w) * sizeof(float);
    if (bytes < 4096) bytes = 4096;   // the MLP CTAs stage 64 x 16 partial sums in the same allocation
    int cap = max_smem_optin();
    if (cap <= 0) cap = 232448;   // no device visible (build container): B200 opt-in limit
    return bytes <= (size_t)cap - 64 ? bytes : 0;
}

// Split path (n_coeff >= 2): the FINEST level runs on the whole GPU with the direct-sum kernels of wavelet.cu -- its adjoint
// with the Adam update of its detail bands in the epilogue, its synthesis at the end -- and only the coarser levels go
// through the per-channel shared-memory kernel.  Three dependent launches instead of six, and the one-SM-per-channel part
// shrinks 8x.  (Measured on B200, C16/G15: the whole pyramid in one CTA per channel takes 29-32 us, issue-bound on 16 of
// 148 SMs; a variant with a cluster of 8 CTAs per channel and the buffers in L2 took 52 us, every pass paying two L2 round
// trips; the six separate launches 24 us.)
namespace lfgc {
int wavelet_finest_bwd_adam(const lfgc_wavelet_desc* w, const float* grad_grid_cl, int Cp, float* low_grad_out, float* p_l,
                            float* g_l, float* m_l, float* v_l, const float* lr, const int* step, const AdamCoef& coef,
                            float w2x2, cudaStream_t st);
int wavelet_finest_fwd(const lfgc_wavelet_desc* w, const float* low_cl, const float* coeff_l, float* grid_cl, int Cp,
                       float* also_zero, cudaStream_t st);
}

static size_t finest_low_elems(const lfgc_wavelet_desc* w) {
    const int l = w->n_coeff - 1;
    return (size_t)w->C * w->dims[l][0] * w->dims[l][1] * w->dims[l][2];
}

extern "C" size_t lfgc_grid_step_scratch_bytes(const lfgc_wavelet_desc* w) {
    if (!w || w->n_coeff < 2 || w->n_coeff > LFGC_MAX_LEVELS) return 0;
    return 2 * finest_low_elems(w) * sizeof(float);
}

// 1: lfgc_grid_step can run this pyramid (whole, or split with the scratch above); 0: use the separate kernels
extern "C" int lfgc_grid_step_supported(const lfgc_wavelet_desc* w) {
    if (!w) return 0;
    if (lfgc_grid_step_smem_bytes(w) > 0) return 1;
    if (w->n_coeff >= 2 && w->n_coeff <= LFGC_MAX_LEVELS) {
        lfgc_wavelet_desc t = *w;
        t.n_coeff = w->n_coeff - 1;
        return lfgc_grid_step_smem_bytes(&t) > 0 ? 1 : 0;
    }
    return 0;
}

static int gstep_fill(gstep::Args& A, const lfgc_wavelet_desc* w, int Cp, const lfgc_grid_step_args* a, int nworkers,
                      bool internal = false) {
    const int rc = gstep_check_desc(w);
    if (rc) return rc;
    if (!a) return fail(LFGC_E_INVALID, "grid_step: null arguments");
    if (!internal && (Cp < w->C || (Cp & 3))) return fail(LFGC_E_INVALID, "Cp=%d must be a multiple of 4 and >= C=%d", Cp, w->C);
    if (a->n_srcs < 1 || a->n_srcs > LFGC_MAX_PEERS) return fail(LFGC_E_UNSUPPORTED, "grid_step: %d gradient sources (1..%d)", a->n_srcs, LFGC_MAX_PEERS);
    if (!a->grid_cl || !a->p || !a->g || !a->m || !a->v || !a->lr || !a->step_count || a->pcount < 0 || a->pstride < a->pcount + 1 || a->nslices < 1)
        return fail(LFGC_E_INVALID, "grid_step: bad buffer arguments");
    gstep_layout(A, w, nworkers);
    A.Cp = Cp;
    A.n_srcs = a->n_srcs;
    for (int r = 0; r < LFGC_MAX_PEERS; ++r) {
        A.grad_grid[r] = r < a->n_srcs ? a->grad_grid[r] : nullptr;
        A.mlp_partials[r] = r < a->n_srcs ? a->mlp_partials[r] : nullptr;
        if (r < a->n_srcs && (!A.grad_grid[r] || (a->pcount > 0 && !A.mlp_partials[r])))
            return fail(LFGC_E_INVALID, "grid_step: gradient source %d is null", r);
    }
    A.nslices = a->nslices;
    A.pstride = a->pstride;
    A.pcount = a->pcount;
    A.zero_grid = a->zero_grid;
    A.grid_cl = a->grid_cl;
    A.p = a->p;
    A.g = a->g;
    A.m = a->m;
    A.v = a->v;
    for (int l = 0; l < gstep::kMaxCoeff; ++l) {
        A.coeff_off[l] = l < w->n_coeff ? a->coeff_off[l] : 0;
        if (l < w->n_coeff && a->coeff_off[l] < 0) return fail(LFGC_E_INVALID, "grid_step: negative coefficient offset");
    }
    A.mlp_off = a->mlp_off;
    A.loss_out = a->loss_out;
    A.lr = a->lr;
    A.step = a->step_count;
    A.c = make_adam_coef(a->beta1, a->beta2, a->eps, a->grad_scale);
    A.w2x2 = (float)(2.0 * a->weight_l2);
    A.n_mlp_ctas = a->pcount > 0 ? (a->pcount + 1 + 63) / 64 : 0;   // 64 parameters per CTA (see the kernel)
    A.panel_image = nullptr;
    if (a->panel_image && a->panel_model) {
        SampleParams P;
        const int prc = fill_sample_params(a->panel_model, 0, P);
        if (prc) return prc;
        A.pm = make_tc_panel_map(P);
        if (A.pm.total_bytes > 0) {
            const int pc = (P.in0 * P.H + P.H) + (P.L - 1) * (P.H * P.H + P.H) + P.H + 1;   // packed MLP block
            if (pc != a->pcount)
                return fail(LFGC_E_INVALID, "grid_step: panel_model has %d MLP parameters, pcount = %d", pc, a->pcount);
            A.panel_image = reinterpret_cast<unsigned char*>(a->panel_image);
        }
    }
    return LFGC_OK;
}

static int gstep_launch(const gstep::Args& A, const lfgc_wavelet_desc* w, cudaStream_t st) {
    const size_t smem = lfgc_grid_step_smem_bytes(w);
    if (smem == 0) return fail(LFGC_E_UNSUPPORTED, "grid_step: the per-channel wavelet pyramid does not fit in shared memory");
    void (*kern)(const gstep::Args) = gstep::grid_step_kernel<4>;
    if (w->n_coeff == 1 || w->n_taps == 2) kern = gstep::grid_step_kernel<2>;
    LFGC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    (void)launch_pdl(kern, dim3((unsigned)(A.C + A.n_mlp_ctas)), dim3(gstep::kThreads), smem, st, A);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

extern "C" int lfgc_grid_step(const lfgc_wavelet_desc* w, int Cp, const lfgc_grid_step_args* a, void* stream) {
    gstep::Args A;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = gstep_check_desc(w);
    if (rc) return rc;
    if (!a) return fail(LFGC_E_INVALID, "grid_step: null arguments");
    // Measured on B200 (profiles/grid_step_timing.py, 32768 samples per step; split / whole pyramid per CTA / separate
    // launches, us per step): two wavelet levels C16/G15 70.6 / 79.5 / 73.7, C32/G15 92.7 / 105.6 / 98.9, C8/G15 62.7 /
    // 65.9 / 67.6, C16/G17 73.7 / 81.7 / 77.8; one level C8/G9 59.4 / 56.0 / 59.3, C16/G9 64.2 / 63.5 / 65.5.
    const bool can_split = w->n_coeff >= 2 && a->n_srcs == 1 && a->scratch && a->scratch_bytes >= lfgc_grid_step_scratch_bytes(w);
    bool split = can_split && (w->n_coeff >= 3 || lfgc_grid_step_smem_bytes(w) == 0);
    if (const char* e = getenv("LFGC_GRID_STEP_SPLIT")) split = can_split && e[0] != '0';   // tuning / test override
    if (!split) {
        rc = gstep_fill(A, w, Cp, a, gstep::kThreads);
        if (rc) return rc;
        return gstep_launch(A, w, st);
    }
    if (Cp < w->C || (Cp & 3)) return fail(LFGC_E_INVALID, "Cp=%d must be a multiple of 4 and >= C=%d", Cp, w->C);
    if (!a->grad_grid[0] || !a->grid_cl || !a->p || !a->g || !a->m || !a->v || !a->lr || !a->step_count)
        return fail(LFGC_E_INVALID, "grid_step: bad buffer arguments");
    const int last = w->n_coeff - 1;
    float* low_grad = a->scratch;
    float* low_val = a->scratch + finest_low_elems(w);
    const long long off = a->coeff_off[last];
    // (1) finest adjoint + Adam of the finest detail bands (reads step_count[0]; the kernel of (2) publishes the increment)
    rc = wavelet_finest_bwd_adam(w, a->grad_grid[0], Cp, low_grad, a->p + off, a->g + off, a->m + off, a->v + off, a->lr,
                                 a->step_count, make_adam_coef(a->beta1, a->beta2, a->eps, a->grad_scale),
                                 (float)(2.0 * a->weight_l2), st);
    if (rc) return rc;
    // (2) the coarser levels, per channel in shared memory, plus the MLP block; channels-last intermediates of stride C
    lfgc_wavelet_desc wt = *w;
    wt.n_coeff = last;
    lfgc_grid_step_args at = *a;
    at.grad_grid[0] = low_grad;
    at.zero_grid = nullptr;
    at.grid_cl = low_val;
    rc = gstep_fill(A, &wt, w->C, &at, gstep::kThreads, true);
    if (rc) return rc;
    rc = gstep_launch(A, &wt, st);
    if (rc) return rc;
    // (3) finest synthesis from the updated low-pass and detail bands; clears the gradient accumulator
    return wavelet_finest_fwd(w, low_val, a->p + off, a->grid_cl, Cp, a->zero_grid, st);
}

#ifdef LFGC_GRID_STEP_HOST_TEST
// Test hook (only in builds made by tests/test_grid_step_host.py, never in liblfgc.so): the SAME per-column functions, the
// same plans and the same phase order, run sequentially on HOST memory, so the separable index arithmetic can be checked
// against the numpy oracle without a GPU.  All pointers are host pointers here.
template <int NT>
static void gstep_run_host(const gstep::Args& A) {
    using namespace gstep;
    using PassT = Pass<NT, false>;
    const size_t total = (size_t)A.sL + A.sX + A.sY + A.sC;
    float* mem = (float*)calloc(total + 4, sizeof(float));
    Bufs S;
    S.L = mem;
    S.X = S.L + A.sL;
    S.Y = S.X + A.sX;
    S.Cg = S.Y + A.sY;
    float flo[NT], fhi[NT];
    for (int i = 0; i < NT; ++i) {
        flo[i] = A.lo[i];
        fhi[i] = A.hi[i];
    }
    const int step = A.step[0] + 1;
    float step_size, bc2_sqrt;
    adam_step_scalars(A.c, step, *A.lr, step_size, bc2_sqrt);
    const int last = A.n_coeff - 1;
    const int nvox = last >= 1 ? A.lv[last].t[0] * A.lv[last].t[1] * A.lv[last].t[2] : A.d0[0] * A.d0[1] * A.d0[2];
    for (int c = 0; c < A.C; ++c) {
        for (int i = 0; i < nvox; ++i) {
            float gsum = 0.0f;
            for (int r = 0; r < A.n_srcs; ++r) gsum += A.grad_grid[r][(long long)i * A.Cp + c];
            S.L[i] = gsum;
        }
        for (int l = last; l >= 1; --l) {
            const LevelPlan& P = A.lv[l];
            for (size_t i = 0; i < (size_t)A.sX; ++i) S.X[i] = 1e30f;   // poison: every element read must have been written
            for (size_t i = 0; i < (size_t)A.sY; ++i) S.Y[i] = 1e30f;
            LFGC_RUN_PASS(az, P.az, 0, 1)
            LFGC_RUN_PASS(ay, P.ay, 0, 1)
            LFGC_RUN_PASS(ax, P.ax, 0, 1)
        }
        for (int l = 0; l <= last; ++l) {
            const int n_l = coeff_elems(A, l);
            float* slots = l == 0 ? S.L : S.Cg + A.lv[l].cg_off;
            const long long base = A.coeff_off[l] + (long long)c * n_l;
            for (int e = 0; e < n_l; ++e) {
                float pi = A.p[base + e], mi = A.m[base + e], vi = A.v[base + e];
                const float gi = fmaf(A.w2x2, pi, slots[e]);
                adam_update(pi, gi, mi, vi, A.c, step_size, bc2_sqrt);
                A.p[base + e] = pi;
                A.m[base + e] = mi;
                A.v[base + e] = vi;
                A.g[base + e] = gi;
                slots[e] = pi;
            }
        }
        for (int l = 1; l <= last; ++l) {
            const LevelPlan& P = A.lv[l];
            for (size_t i = 0; i < (size_t)A.sX; ++i) S.X[i] = 1e30f;
            for (size_t i = 0; i < (size_t)A.sY; ++i) S.Y[i] = 1e30f;
            LFGC_RUN_PASS(sx, P.sx, 0, 1)
            LFGC_RUN_PASS(sy, P.sy, 0, 1)
            LFGC_RUN_PASS(sz, P.sz, 0, 1)
        }
        for (int i = 0; i < nvox; ++i) A.grid_cl[(long long)i * A.Cp + c] = S.L[i];
    }
    for (int i = 0; i < nvox; ++i)
        for (int c = A.C; c < A.Cp; ++c) A.grid_cl[(long long)i * A.Cp + c] = 0.0f;
    if (A.zero_grid)
        for (long long i = 0; i < (long long)nvox * A.Cp; ++i) A.zero_grid[i] = 0.0f;
    for (int j = 0; j <= A.pcount; ++j) {
        float t = 0.0f;
        for (int r = 0; r < A.n_srcs; ++r)
            for (int b = 0; b < A.nslices; ++b) t += A.mlp_partials[r][(size_t)b * A.pstride + j];
        if (j < A.pcount) {
            const long long i = A.mlp_off + j;
            float pi = A.p[i], mi = A.m[i], vi = A.v[i];
            adam_update(pi, t, mi, vi, A.c, step_size, bc2_sqrt);
            A.p[i] = pi;
            A.m[i] = mi;
            A.v[i] = vi;
            A.g[i] = t;
        } else if (A.loss_out) {
            A.loss_out[0] = t;
        }
    }
    A.step[0] = step;
    free(mem);
}

extern "C" int lfgc_grid_step_host(const lfgc_wavelet_desc* w, int Cp, const lfgc_grid_step_args* a, int nworkers) {
    gstep::Args A;
    const int rc = gstep_fill(A, w, Cp, a, nworkers);
    if (rc) return rc;
    if (w->n_coeff == 1 || w->n_taps == 2) gstep_run_host<2>(A);
    else gstep_run_host<4>(A);
    return LFGC_OK;
}
#endif
