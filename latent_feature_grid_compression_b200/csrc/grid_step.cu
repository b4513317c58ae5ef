// "Grid step" kernel: everything of one optimiser step that is NOT per-sample, for mask-free models, in ONE launch:
//
//     [sum of the data-parallel ranks' gradient buffers, read over NVLink from peer memory]
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
#include "lfgc_common.cuh"

#include <stdlib.h>

namespace lfgc {
namespace gstep {

constexpr int kThreads = 1024;

struct Args {
    int n_coeff, C, Cp, ntaps;
    float lo[LFGC_MAX_TAPS], hi[LFGC_MAX_TAPS];
    int d[LFGC_MAX_LEVELS][3], t[LFGC_MAX_LEVELS][3], off[LFGC_MAX_LEVELS][3];
    int cg_off[LFGC_MAX_LEVELS];      // offset (floats) of level l's detail bands inside the shared coefficient block
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
    long long coeff_off[LFGC_MAX_LEVELS];
    long long mlp_off;
    float* loss_out;
    const float* lr;
    int* step;
    AdamCoef c;
    float w2x2;                       // 2 * weight_l2
    int n_mlp_ctas;
    int rank;                         // data-parallel peer barrier (sync_epoch != nullptr): this rank's index, ...
    int* sync_flags[LFGC_MAX_PEERS];  // ... every rank's flag array (int[n_srcs], peer memory; [rank] is the local one)
    int* sync_epoch;                  // ... and this rank's launch counter (device int, monotonic)
};

// Shared-memory (or, in the host test, heap) partitions of one channel
struct Bufs {
    float* L;    // current low-pass volume / grid gradient / coefficient tensor 0 of this channel
    float* X;    // x-pass intermediate: [4 (a,b)][d0][d1][t2]
    float* Y;    // y-pass intermediate: [2 (a)][d0][t1][t2]
    float* Cg;   // detail bands of every level: gradient after the adjoint, updated value after Adam
    const float* lo;   // reconstruction taps (shared-memory copies in the kernel: the tap index differs between the
    const float* hi;   // lanes of a warp, which would serialise constant-bank reads)
};

// Every pass is organised by COLUMNS: a work unit is one 1-D line of the pass (fixed position in the two untouched
// dimensions) times a range [o0, o1) of output positions along the filtered dimension.  The index decomposition happens
// once per unit, the inner loops are division-free and the filter-tap selection is uniform over a warp.  (A first version
// decomposed a flat index per OUTPUT ELEMENT: six integer divisions per element made the kernel 44 us, slower than the
// six launches it replaces.)

// ---- synthesis level l, separable: x pass -> y pass -> z pass ----------------------------------------------------------
// band k = 4a + 2b + c, (a, b, c) = filter along (dim0, dim1, dim2), 0 = low / 1 = high; k = 0 is the running low-pass
// (Torch_Wavelet_Transform.py:39-57,91-104).  out[o] = sum_i in[i] * f[o + off - 2 i].
// x pass: column = (ab, iz, iy), 4 d0 d1 of them; outputs ox in [o0, o1) of t2
__host__ __device__ __forceinline__ void synth_x(const Args& A, const Bufs& S, int l, int col, int o0, int o1) {
    const int d0 = A.d[l][0], d1 = A.d[l][1], d2 = A.d[l][2], t2 = A.t[l][2];
    const int dvol = d0 * d1 * d2;
    const int ab = col / (d0 * d1);
    const int zy = col - ab * d0 * d1;   // iz * d1 + iy
    const float* b0 = (ab == 0 ? S.L : S.Cg + A.cg_off[l] + (2 * ab - 1) * dvol) + zy * d2;   // band (a, b, 0)
    const float* b1 = S.Cg + A.cg_off[l] + (2 * ab) * dvol + zy * d2;                         // band (a, b, 1)
    float* out = S.X + col * t2;
    for (int ox = o0; ox < o1; ++ox) {
        const int o = ox + A.off[l][2];
        float acc = 0.0f;
        for (int a = 0; 2 * a < A.ntaps; ++a) {
            const int i = (o >> 1) - a;
            const int tt = (o & 1) + 2 * a;
            if (i >= 0 && i < d2) {
                acc = fmaf(b0[i], S.lo[tt], acc);
                acc = fmaf(b1[i], S.hi[tt], acc);
            }
        }
        out[ox] = acc;
    }
}
// y pass: column = (a, iz, ox), 2 d0 t2 of them; outputs oy in [o0, o1) of t1
__host__ __device__ __forceinline__ void synth_y(const Args& A, const Bufs& S, int l, int col, int o0, int o1) {
    const int d0 = A.d[l][0], d1 = A.d[l][1], t1 = A.t[l][1], t2 = A.t[l][2];
    const int az = col / t2;             // a * d0 + iz
    const int ox = col - az * t2;
    const int a_ = az / d0;
    const int iz = az - a_ * d0;
    const int plane = d0 * d1 * t2;
    const float* x0 = S.X + (2 * a_) * plane + iz * d1 * t2 + ox;       // (a, b = 0)
    const float* x1 = x0 + plane;                                        // (a, b = 1)
    float* out = S.Y + az * t1 * t2 + ox;
    for (int oy = o0; oy < o1; ++oy) {
        const int o = oy + A.off[l][1];
        float acc = 0.0f;
        for (int a = 0; 2 * a < A.ntaps; ++a) {
            const int i = (o >> 1) - a;
            const int tt = (o & 1) + 2 * a;
            if (i >= 0 && i < d1) {
                acc = fmaf(x0[i * t2], S.lo[tt], acc);
                acc = fmaf(x1[i * t2], S.hi[tt], acc);
            }
        }
        out[oy * t2] = acc;
    }
}
// z pass: column = (oy, ox), t1 t2 of them; outputs oz in [o0, o1) of t0
__host__ __device__ __forceinline__ void synth_z(const Args& A, const Bufs& S, int l, int col, int o0, int o1) {
    const int d0 = A.d[l][0], t1 = A.t[l][1], t2 = A.t[l][2];
    const int pl = t1 * t2;
    const float* y0 = S.Y + col;
    const float* y1 = y0 + d0 * pl;
    for (int oz = o0; oz < o1; ++oz) {
        const int o = oz + A.off[l][0];
        float acc = 0.0f;
        for (int a = 0; 2 * a < A.ntaps; ++a) {
            const int i = (o >> 1) - a;
            const int tt = (o & 1) + 2 * a;
            if (i >= 0 && i < d0) {
                acc = fmaf(y0[i * pl], S.lo[tt], acc);
                acc = fmaf(y1[i * pl], S.hi[tt], acc);
            }
        }
        S.L[oz * pl + col] = acc;
    }
}

// ---- adjoint of synthesis level l, separable: z^T pass -> y^T pass -> x^T pass ------------------------------------------
// g_in[i] = sum_tt g_out[2 i + tt - off] * f[tt]
// z^T pass: column = (oy, ox), t1 t2 of them; outputs (a, iz) = o in [o0, o1) of 2 d0
__host__ __device__ __forceinline__ void adj_z(const Args& A, const Bufs& S, int l, int col, int o0, int o1) {
    const int d0 = A.d[l][0], t0 = A.t[l][0], t1 = A.t[l][1], t2 = A.t[l][2];
    const int pl = t1 * t2;
    for (int o = o0; o < o1; ++o) {
        const int a_ = o >= d0;
        const int iz = o - a_ * d0;
        const float* f = a_ ? S.hi : S.lo;
        float acc = 0.0f;
        for (int tt = 0; tt < A.ntaps; ++tt) {
            const int q = 2 * iz + tt - A.off[l][0];
            if (q >= 0 && q < t0) acc = fmaf(S.L[q * pl + col], f[tt], acc);
        }
        S.Y[o * pl + col] = acc;
    }
}
// y^T pass: column = (a, iz, ox), 2 d0 t2 of them; outputs (b, iy) = o in [o0, o1) of 2 d1
__host__ __device__ __forceinline__ void adj_y(const Args& A, const Bufs& S, int l, int col, int o0, int o1) {
    const int d0 = A.d[l][0], d1 = A.d[l][1], t1 = A.t[l][1], t2 = A.t[l][2];
    const int az = col / t2;             // a * d0 + iz
    const int ox = col - az * t2;
    const int a_ = az / d0;
    const int iz = az - a_ * d0;
    const float* y = S.Y + az * t1 * t2 + ox;
    for (int o = o0; o < o1; ++o) {
        const int b_ = o >= d1;
        const int iy = o - b_ * d1;
        const float* f = b_ ? S.hi : S.lo;
        float acc = 0.0f;
        for (int tt = 0; tt < A.ntaps; ++tt) {
            const int q = 2 * iy + tt - A.off[l][1];
            if (q >= 0 && q < t1) acc = fmaf(y[q * t2], f[tt], acc);
        }
        S.X[(((2 * a_ + b_) * d0 + iz) * d1 + iy) * t2 + ox] = acc;
    }
}
// x^T pass: column = (ab, iz, iy), 4 d0 d1 of them; outputs (c, ix) = o in [o0, o1) of 2 d2
__host__ __device__ __forceinline__ void adj_x(const Args& A, const Bufs& S, int l, int col, int o0, int o1) {
    const int d0 = A.d[l][0], d1 = A.d[l][1], d2 = A.d[l][2], t2 = A.t[l][2];
    const int dvol = d0 * d1 * d2;
    const int ab = col / (d0 * d1);
    const int zy = col - ab * d0 * d1;
    const float* x = S.X + col * t2;
    for (int o = o0; o < o1; ++o) {
        const int c_ = o >= d2;
        const int ix = o - c_ * d2;
        const float* f = c_ ? S.hi : S.lo;
        float acc = 0.0f;
        for (int tt = 0; tt < A.ntaps; ++tt) {
            const int q = 2 * ix + tt - A.off[l][2];
            if (q >= 0 && q < t2) acc = fmaf(x[q], f[tt], acc);
        }
        const int k = 2 * ab + c_;
        const int pos = zy * d2 + ix;
        if (k == 0) S.L[pos] = acc;                                   // gradient of the low-pass input of this level
        else S.Cg[A.cg_off[l] + (k - 1) * dvol + pos] = acc;          // gradient of detail band k
    }
}

// Work distribution of one pass over `nthreads` workers: ncols columns x n_out outputs; when there are fewer columns
// than workers the output range is cut into groups so that (almost) every worker has something to do.
struct PassPlan {
    int ncols, n_out, groups, per;
};
__host__ __device__ __forceinline__ PassPlan plan_pass(int ncols, int n_out, int nthreads) {
    PassPlan p;
    p.ncols = ncols;
    p.n_out = n_out;
    int g = ncols > 0 ? nthreads / ncols : 1;
    if (g < 1) g = 1;
    if (g > n_out) g = n_out;
    p.groups = g;
    p.per = (n_out + g - 1) / g;
    return p;
}
#define LFGC_PASS(fn, ncols_, nout_)                                                           \
    {                                                                                          \
        const PassPlan pp = plan_pass((ncols_), (nout_), kThreads);                            \
        for (int u = tid; u < pp.ncols * pp.groups; u += kThreads) {                           \
            const int grp = u / pp.ncols;                                                      \
            const int col = u - grp * pp.ncols;                                                \
            const int b0_ = grp * pp.per;                                                      \
            const int b1_ = b0_ + pp.per < pp.n_out ? b0_ + pp.per : pp.n_out;                 \
            fn(A, S, l, col, b0_, b1_);                                                        \
        }                                                                                      \
        __syncthreads();                                                                       \
    }

// Adam on element e of coefficient tensor l of channel c; the gradient sits in shared memory and is replaced there by
// the updated coefficient (what the synthesis below reads)
__host__ __device__ __forceinline__ void adam_coeff(const Args& A, const Bufs& S, int l, int c, int e, int n_l,
                                                    float step_size, float bc2_sqrt) {
    float* slot = l == 0 ? S.L + e : S.Cg + A.cg_off[l] + e;
    const long long i = A.coeff_off[l] + (long long)c * n_l + e;
    float pi = A.p[i], mi = A.m[i], vi = A.v[i];
    const float gi = fmaf(A.w2x2, pi, *slot);
    adam_update(pi, gi, mi, vi, A.c, step_size, bc2_sqrt);
    A.p[i] = pi;
    A.m[i] = mi;
    A.v[i] = vi;
    A.g[i] = gi;
    *slot = pi;
}

__host__ __device__ __forceinline__ int coeff_elems(const Args& A, int l) {
    const int dv = A.d[l][0] * A.d[l][1] * A.d[l][2];
    return l == 0 ? dv : 7 * dv;
}

#ifdef __CUDACC__
// Data-parallel barrier INSIDE the kernel (no separate collective, no extra launch): every rank's kernel announces its
// epoch to all ranks' flag arrays with a system-scope release store and waits until every rank has announced the same
// epoch.  A rank reaches this point only after its own per-sample kernel (same stream) has completed, so once the wait
// is over every rank's gradient buffer of this step is complete and visible over NVLink.  Epochs only grow, so there is
// nothing to reset; a lost peer traps after ~2 s instead of hanging the device.
__device__ __forceinline__ void peer_barrier(const Args& A, int tid) {
    const int e = *reinterpret_cast<volatile int*>(A.sync_epoch) + 1;
    if (tid < A.n_srcs) {
        if (blockIdx.x == 0) {
            __threadfence_system();
            asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(A.sync_flags[tid] + A.rank), "r"(e) : "memory");
        }
        const int* mine = A.sync_flags[A.rank] + tid;
        const long long t0 = clock64();
        for (;;) {
            int seen;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
            if (seen >= e) break;
            if (clock64() - t0 > 4000000000ll) __trap();
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 1) grid_step_kernel(const __grid_constant__ Args A) {
    LFGC_PDL_PROLOGUE();
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_step_size, s_bc2_sqrt;
    __shared__ int s_step;
    __shared__ float s_lo[LFGC_MAX_TAPS], s_hi[LFGC_MAX_TAPS];
    const int tid = threadIdx.x;
    if (A.sync_epoch) peer_barrier(A, tid);
    if (tid < LFGC_MAX_TAPS) {
        s_lo[tid] = A.lo[tid];
        s_hi[tid] = A.hi[tid];
    }
    if (tid == 0) {
        const int step = *reinterpret_cast<volatile int*>(A.step) + 1;
        float step_size, bc2_sqrt;
        adam_step_scalars(A.c, step, *A.lr, step_size, bc2_sqrt);
        s_step = step;
        s_step_size = step_size;
        s_bc2_sqrt = bc2_sqrt;
    }
    const int c = blockIdx.x;
    if (c < A.C) {
        Bufs S;
        S.L = smem;
        S.X = S.L + A.sL;
        S.Y = S.X + A.sX;
        S.Cg = S.Y + A.sY;
        S.lo = s_lo;
        S.hi = s_hi;
        const int last = A.n_coeff - 1;
        // ---- this channel's slice of the grid gradient (summed over the ranks in rank order), accumulator cleared ----------
        const int* G = last >= 1 ? A.t[last] : A.d[0];
        const int nvox = G[0] * G[1] * G[2];
        for (int i = tid; i < nvox; i += kThreads) {
            const long long a = (long long)i * A.Cp + c;
            float gsum = __ldcv(A.grad_grid[0] + a);          // rewritten every step (also by peers): never from L1
            for (int r = 1; r < A.n_srcs; ++r) gsum += __ldcv(A.grad_grid[r] + a);
            S.L[i] = gsum;
            if (A.zero_grid) A.zero_grid[a] = 0.0f;
        }
        if (A.zero_grid && c == A.C - 1) {                      // pad channels of the accumulator stay zero anyway; keep them so
            for (int i = tid; i < nvox * (A.Cp - A.C); i += kThreads)
                A.zero_grid[(long long)(i / (A.Cp - A.C)) * A.Cp + A.C + i % (A.Cp - A.C)] = 0.0f;
        }
        __syncthreads();
        // ---- adjoint, finest -> coarsest --------------------------------------------------------------------------------------
        for (int l = last; l >= 1; --l) {
            const int d0 = A.d[l][0], d1 = A.d[l][1], d2 = A.d[l][2], t1 = A.t[l][1], t2 = A.t[l][2];
            LFGC_PASS(adj_z, t1 * t2, 2 * d0)
            LFGC_PASS(adj_y, 2 * d0 * t2, 2 * d1)
            LFGC_PASS(adj_x, 4 * d0 * d1, 2 * d2)
        }
        // ---- Adam on this channel's coefficients ---------------------------------------------------------------------------------
        const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
        for (int l = 0; l <= last; ++l) {
            const int n_l = coeff_elems(A, l);
            float* slots = l == 0 ? S.L : S.Cg + A.cg_off[l];
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
            const int d0 = A.d[l][0], d1 = A.d[l][1], t0 = A.t[l][0], t1 = A.t[l][1], t2 = A.t[l][2];
            LFGC_PASS(synth_x, 4 * d0 * d1, t2)
            LFGC_PASS(synth_y, 2 * d0 * t2, t1)
            LFGC_PASS(synth_z, t1 * t2, t0)
        }
        for (int i = tid; i < nvox; i += kThreads) A.grid_cl[(long long)i * A.Cp + c] = S.L[i];
        if (c == A.C - 1) {
            for (int i = tid; i < nvox * (A.Cp - A.C); i += kThreads)
                A.grid_cl[(long long)(i / (A.Cp - A.C)) * A.Cp + A.C + i % (A.Cp - A.C)] = 0.0f;
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
            if (A.sync_epoch) *A.sync_epoch += 1;
            __threadfence();
            A.step[0] = s_step;
        }
    }
}
#endif

}  // namespace gstep
}  // namespace lfgc

using namespace lfgc;

// Shared-memory partition sizes; returns the total in floats (0: no coefficient tensor)
static size_t gstep_layout(gstep::Args& A, const lfgc_wavelet_desc* w) {
    A.n_coeff = w->n_coeff;
    A.C = w->C;
    A.ntaps = w->n_taps;
    for (int i = 0; i < LFGC_MAX_TAPS; ++i) {
        A.lo[i] = i < w->n_taps ? w->rec_lo[i] : 0.0f;
        A.hi[i] = i < w->n_taps ? w->rec_hi[i] : 0.0f;
    }
    size_t sL = 0, sX = 0, sY = 0, sC = 0;
    for (int l = 0; l < LFGC_MAX_LEVELS; ++l) {
        for (int a = 0; a < 3; ++a) {
            A.d[l][a] = l < w->n_coeff ? w->dims[l][a] : 0;
            A.t[l][a] = (l >= 1 && l < w->n_coeff) ? w->target[l][a] : 0;
            A.off[l][a] = (l >= 1 && l < w->n_coeff) ? (2 * A.d[l][a] + A.ntaps - 2 - A.t[l][a]) / 2 : 0;
        }
        A.cg_off[l] = 0;
        if (l >= w->n_coeff) continue;
        const size_t dv = (size_t)A.d[l][0] * A.d[l][1] * A.d[l][2];
        if (dv > sL) sL = dv;
        if (l >= 1) {
            const size_t tv = (size_t)A.t[l][0] * A.t[l][1] * A.t[l][2];
            if (tv > sL) sL = tv;
            const size_t x = 4 * (size_t)A.d[l][0] * A.d[l][1] * A.t[l][2];
            const size_t y = 2 * (size_t)A.d[l][0] * A.t[l][1] * A.t[l][2];
            if (x > sX) sX = x;
            if (y > sY) sY = y;
            A.cg_off[l] = (int)sC;
            sC += 7 * dv;
        }
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
    if (w->n_coeff > 1 && (w->n_taps < 2 || w->n_taps > LFGC_MAX_TAPS || (w->n_taps & 1)))
        return fail(LFGC_E_INVALID, "grid_step: bad filter length %d", w->n_taps);
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

static int gstep_fill(gstep::Args& A, const lfgc_wavelet_desc* w, int Cp, const lfgc_grid_step_args* a) {
    const int rc = gstep_check_desc(w);
    if (rc) return rc;
    if (!a) return fail(LFGC_E_INVALID, "grid_step: null arguments");
    if (Cp < w->C || (Cp & 3)) return fail(LFGC_E_INVALID, "Cp=%d must be a multiple of 4 and >= C=%d", Cp, w->C);
    if (a->n_srcs < 1 || a->n_srcs > LFGC_MAX_PEERS) return fail(LFGC_E_UNSUPPORTED, "grid_step: %d gradient sources (1..%d)", a->n_srcs, LFGC_MAX_PEERS);
    if (!a->grid_cl || !a->p || !a->g || !a->m || !a->v || !a->lr || !a->step_count || a->pcount < 0 || a->pstride < a->pcount + 1 || a->nslices < 1)
        return fail(LFGC_E_INVALID, "grid_step: bad buffer arguments");
    gstep_layout(A, w);
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
    for (int l = 0; l < LFGC_MAX_LEVELS; ++l) {
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
    A.rank = a->rank;
    A.sync_epoch = a->sync_epoch;
    for (int r = 0; r < LFGC_MAX_PEERS; ++r) {
        A.sync_flags[r] = (a->sync_epoch && r < a->n_srcs) ? a->sync_flags[r] : nullptr;
        if (a->sync_epoch && r < a->n_srcs && !A.sync_flags[r]) return fail(LFGC_E_INVALID, "grid_step: flag array of rank %d is null", r);
    }
    if (a->sync_epoch && (a->rank < 0 || a->rank >= a->n_srcs)) return fail(LFGC_E_INVALID, "grid_step: bad rank %d", a->rank);
    return LFGC_OK;
}

extern "C" int lfgc_grid_step(const lfgc_wavelet_desc* w, int Cp, const lfgc_grid_step_args* a, void* stream) {
    gstep::Args A;
    const int rc = gstep_fill(A, w, Cp, a);
    if (rc) return rc;
    const size_t smem = lfgc_grid_step_smem_bytes(w);
    if (smem == 0) return fail(LFGC_E_UNSUPPORTED, "grid_step: the per-channel wavelet pyramid does not fit in shared memory");
    auto kern = gstep::grid_step_kernel;
    LFGC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    (void)launch_pdl(kern, dim3((unsigned)(A.C + A.n_mlp_ctas)), dim3(gstep::kThreads), smem, (cudaStream_t)stream, A);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

#ifdef LFGC_GRID_STEP_HOST_TEST
// Test hook (only in builds made by tests/test_grid_step_host.py, never in liblfgc.so): the SAME per-element functions in
// the same phase order, run sequentially on HOST memory, so the separable index arithmetic can be checked against the
// numpy oracle without a GPU.  All pointers are host pointers here.
#define HOST_PASS(fn, ncols_, nout_)                                                           \
    {                                                                                          \
        const PassPlan pp = plan_pass((ncols_), (nout_), kThreads);                            \
        for (int u = 0; u < pp.ncols * pp.groups; ++u) {                                       \
            const int grp = u / pp.ncols, col = u - grp * pp.ncols;                            \
            const int b0_ = grp * pp.per;                                                      \
            const int b1_ = b0_ + pp.per < pp.n_out ? b0_ + pp.per : pp.n_out;                 \
            fn(A, S, l, col, b0_, b1_);                                                        \
        }                                                                                      \
    }
extern "C" int lfgc_grid_step_host(const lfgc_wavelet_desc* w, int Cp, const lfgc_grid_step_args* a) {
    using namespace gstep;
    Args A;
    const int rc = gstep_fill(A, w, Cp, a);
    if (rc) return rc;
    const size_t total = (size_t)A.sL + A.sX + A.sY + A.sC;
    float* mem = (float*)calloc(total + 4, sizeof(float));
    Bufs S;
    S.L = mem;
    S.X = S.L + A.sL;
    S.Y = S.X + A.sX;
    S.Cg = S.Y + A.sY;
    S.lo = A.lo;
    S.hi = A.hi;
    const int step = A.step[0] + 1;
    float step_size, bc2_sqrt;
    adam_step_scalars(A.c, step, *A.lr, step_size, bc2_sqrt);
    const int last = A.n_coeff - 1;
    const int* G = last >= 1 ? A.t[last] : A.d[0];
    const int nvox = G[0] * G[1] * G[2];
    for (int c = 0; c < A.C; ++c) {
        for (int i = 0; i < nvox; ++i) {
            float gsum = 0.0f;
            for (int r = 0; r < A.n_srcs; ++r) gsum += A.grad_grid[r][(long long)i * A.Cp + c];
            S.L[i] = gsum;
        }
        for (int l = last; l >= 1; --l) {
            const int d0 = A.d[l][0], d1 = A.d[l][1], d2 = A.d[l][2], t1 = A.t[l][1], t2 = A.t[l][2];
            // same column / group decomposition as the kernel (plan_pass with the kernel's thread count)
            HOST_PASS(adj_z, t1 * t2, 2 * d0)
            HOST_PASS(adj_y, 2 * d0 * t2, 2 * d1)
            HOST_PASS(adj_x, 4 * d0 * d1, 2 * d2)
        }
        for (int l = 0; l <= last; ++l) {
            const int n_l = coeff_elems(A, l);
            for (int e = 0; e < n_l; ++e) adam_coeff(A, S, l, c, e, n_l, step_size, bc2_sqrt);
        }
        for (int l = 1; l <= last; ++l) {
            const int d0 = A.d[l][0], d1 = A.d[l][1], t0 = A.t[l][0], t1 = A.t[l][1], t2 = A.t[l][2];
            HOST_PASS(synth_x, 4 * d0 * d1, t2)
            HOST_PASS(synth_y, 2 * d0 * t2, t1)
            HOST_PASS(synth_z, t1 * t2, t0)
        }
        for (int i = 0; i < nvox; ++i) A.grid_cl[(long long)i * A.Cp + c] = S.L[i];
    }
    for (int i = 0; i < nvox; ++i)
        for (int c = A.C; c < A.Cp; ++c) A.grid_cl[(long long)i * A.Cp + c] = 0.0f;
    if (A.zero_grid)
        for (long long i = 0; i < (long long)nvox * A.Cp; ++i) A.zero_grid[i] = 0.0f;
    const int ns = A.nslices;
    for (int j = 0; j <= A.pcount; ++j) {
        float t = 0.0f;
        for (int r = 0; r < A.n_srcs; ++r)
            for (int b = 0; b < ns; ++b) t += A.mlp_partials[r][(size_t)b * A.pstride + j];
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
    return LFGC_OK;
}
#endif
