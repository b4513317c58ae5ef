// Mask layers + multi-level 3-D inverse wavelet synthesis (and its adjoint).
//
// Replaces, per training step / per reconstruction:
//   model/Feature_Grid_Model.py:102-108  decode_volume (mask layer on the coefficients, cat, filter.decode per level)
//   wavelet_transform/Torch_Wavelet_Transform.py:91-104  grouped conv_transpose3d stride 2 + crop
//   model/{Smallify_Dropout,Variational_Dropout_Layer,Straight_Through_Dropout}.py forward of the mask layers
// and their autograd backward.  The decoded grid is emitted channels-last ([z][y][x][Cp]) so that the gather in
// the sample kernels reads one corner as contiguous 16-byte vectors.
#include "lfgc_common.cuh"

namespace lfgc {

// ---------------------------------------------------------------------------------------------------------------
// mask layers
// ---------------------------------------------------------------------------------------------------------------
__global__ void mask_multiplier_kernel(int mode, int64_t n, const float* __restrict__ p0,
                                       const float* __restrict__ p1, const float* __restrict__ noise,
                                       float threshold, float* __restrict__ mult, float* __restrict__ aux) {
    LFGC_PDL_PROLOGUE();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m, a;
    if (mode == LFGC_MASK_DIRECT) {
        m = a = p0[i];
    } else if (mode == LFGC_MASK_VARIATIONAL) {
        // w = exp(log_thetas) + exp(log_var / 2) * xi      (Variational_Dropout_Layer.py:104-108)
        float th = expf(p0[i]);
        float sg = expf(p1[i] / 2.0f);
        m = a = __fadd_rn(th, __fmul_rn(sg, noise[i]));
    } else if (mode == LFGC_MASK_STE_SIGMOID) {
        // value x*[sigmoid(v) >= t], gradient of x*sigmoid(v)  (Straight_Through_Dropout.py:55-58)
        float s = 1.0f / (1.0f + expf(-p0[i]));
        m = (s >= threshold) ? 1.0f : 0.0f;
        a = s;
    } else if (mode == LFGC_MASK_BERNOULLI) {
        m = a = (noise[i] < p0[i]) ? 1.0f : 0.0f;  // STEFunction.forward: input < thresh (:12-13)
    } else {
        m = a = 1.0f;
    }
    mult[i] = m;
    if (aux) aux[i] = a;
}

__global__ void mask_param_grad_kernel(int mode, int64_t n, const float* __restrict__ p0,
                                       const float* __restrict__ p1, const float* __restrict__ noise,
                                       const float* __restrict__ gmult, float* __restrict__ g0,
                                       float* __restrict__ g1, int accumulate) {
    LFGC_PDL_PROLOGUE();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gm = gmult[i];
    if (mode == LFGC_MASK_DIRECT) {
        g0[i] = accumulate ? g0[i] + gm : gm;
    } else if (mode == LFGC_MASK_VARIATIONAL) {
        float th = expf(p0[i]);
        float sg = expf(p1[i] / 2.0f);
        float a = gm * th;
        float b = gm * (0.5f * sg * noise[i]);
        g0[i] = accumulate ? g0[i] + a : a;
        g1[i] = accumulate ? g1[i] + b : b;
    } else if (mode == LFGC_MASK_STE_SIGMOID) {
        float s = 1.0f / (1.0f + expf(-p0[i]));
        float a = gm * (s * (1.0f - s));
        g0[i] = accumulate ? g0[i] + a : a;
    }
}

__global__ void smallify_ema_kernel(const float* __restrict__ betas, float* __restrict__ ema,
                                    float* __restrict__ emavar, int64_t n, float momentum) {
    LFGC_PDL_PROLOGUE();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float b = betas[i];
    float sgn = (b > 0.0f) ? 1.0f : ((b < 0.0f) ? -1.0f : 0.0f);
    float e = ema[i];
    float phi = __fsub_rn(sgn, e);
    // EMA = EMA + m*phi ; EMAVar = (1 - m) * (EMAVar + m * phi^2)      (Smallify_Dropout.py:108-112)
    ema[i] = __fadd_rn(e, __fmul_rn(momentum, phi));
    float one_m = __fsub_rn(1.0f, momentum);
    emavar[i] = __fmul_rn(one_m, __fadd_rn(emavar[i], __fmul_rn(momentum, __fmul_rn(phi, phi))));
}

// ---------------------------------------------------------------------------------------------------------------
// one synthesis level
// ---------------------------------------------------------------------------------------------------------------
// Layouts: coefficient tensors (the nn.Parameters) stay in the reference layout, channels-first (C,[7,]d0,d1,d2);
// everything the kernels produce (intermediate low-pass volumes, the final grid, their gradients) is channels-last
// with the channel as the fastest thread index, so that neighbouring lanes read/write neighbouring addresses.
struct LevelArgs {
    const float* low;    // coarsest level: coeff[0] (C,d0,d1,d2) channels-first; otherwise (d0,d1,d2,C) channels-last
    const float* high;   // (C, 7, d0, d1, d2)
    const float* mlow;   // (d0, d1, d2) or null: multiplier of the low-pass input (coarsest level only)
    const float* mhigh;  // (7, d0, d1, d2) or null
    float* out;          // (t0, t1, t2, Cs) channels-last, Cs = Cp for the final grid, C otherwise
    float* also_zero;    // optional buffer of the same shape as `out`, cleared in the same pass (gradient accumulator)
    int C, Cs, low_cl;
    int d[3], t[3], off[3];
    int ntaps;
    float lo[LFGC_MAX_TAPS], hi[LFGC_MAX_TAPS];
};

// u[o] = sum_k sum_i c_k[i] r_a[oz-2iz] r_b[oy-2iy] r_c[ox-2ix], out[p] = u[p + off], sub-band k = 4a+2b+c
// (Torch_Wavelet_Transform.py:39-57 outer-product filter bank; :100-104 transposed conv + crop).
// With NT taps an output position sees NT/2 input positions per axis: i = (o >> 1) - a, tap (o & 1) + 2a.  The
// template unrolls all (NT/2)^3 x 8 coefficient loads so that they are issued back to back (one L2 round trip)
// instead of one dependent load per loop iteration; out-of-range positions are clamped and get weight zero.
template <int NT>
__global__ void idwt_level_kernel(LevelArgs A) {
    LFGC_PDL_PROLOGUE();
    constexpr int NP = NT / 2;
    const int64_t nvox = (int64_t)A.t[0] * A.t[1] * A.t[2];
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nvox * A.Cs) return;
    const int c = (int)(idx % A.Cs);
    const int64_t p = idx / A.Cs;
    if (A.also_zero) A.also_zero[idx] = 0.0f;
    if (c >= A.C) {
        A.out[idx] = 0.0f;  // zero pad channels
        return;
    }
    const int px = (int)(p % A.t[2]);
    const int py = (int)((p / A.t[2]) % A.t[1]);
    const int pz = (int)(p / ((int64_t)A.t[2] * A.t[1]));
    const int o[3] = {pz + A.off[0], py + A.off[1], px + A.off[2]};
    const int64_t dvol = (int64_t)A.d[0] * A.d[1] * A.d[2];
    const float* highc = A.high + (int64_t)c * 7 * dvol;

    int ii[3][NP];
    float wl[3][NP], wh[3][NP];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax)
#pragma unroll
        for (int a = 0; a < NP; ++a) {
            const int i = (o[ax] >> 1) - a;
            const int t = (o[ax] & 1) + 2 * a;
            const bool ok = i >= 0 && i < A.d[ax];
            ii[ax][a] = ok ? i : 0;
            wl[ax][a] = ok ? A.lo[t] : 0.0f;
            wh[ax][a] = ok ? A.hi[t] : 0.0f;
        }
    float acc = 0.0f;
#pragma unroll
    for (int az = 0; az < NP; ++az)
#pragma unroll
        for (int ay = 0; ay < NP; ++ay)
#pragma unroll
            for (int axx = 0; axx < NP; ++axx) {
                const int64_t b = ((int64_t)ii[0][az] * A.d[1] + ii[1][ay]) * A.d[2] + ii[2][axx];
                float v[8];
                v[0] = A.low_cl ? A.low[b * A.C + c] : A.low[(int64_t)c * dvol + b];
                if (A.mlow) v[0] *= A.mlow[b];
#pragma unroll
                for (int k = 1; k < 8; ++k) {
                    v[k] = highc[(int64_t)(k - 1) * dvol + b];
                    if (A.mhigh) v[k] *= A.mhigh[(int64_t)(k - 1) * dvol + b];
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float wz = (k & 4) ? wh[0][az] : wl[0][az];
                    const float wy = (k & 2) ? wh[1][ay] : wl[1][ay];
                    const float wx = (k & 1) ? wh[2][axx] : wl[2][axx];
                    acc = fmaf(v[k], wz * wy * wx, acc);
                }
            }
    A.out[idx] = acc;
}

// any (even) filter length: plain loops
__global__ void idwt_level_generic_kernel(LevelArgs A) {
    LFGC_PDL_PROLOGUE();
    const int64_t nvox = (int64_t)A.t[0] * A.t[1] * A.t[2];
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nvox * A.Cs) return;
    const int c = (int)(idx % A.Cs);
    const int64_t p = idx / A.Cs;
    if (A.also_zero) A.also_zero[idx] = 0.0f;
    if (c >= A.C) {
        A.out[idx] = 0.0f;
        return;
    }
    int px = (int)(p % A.t[2]);
    int py = (int)((p / A.t[2]) % A.t[1]);
    int pz = (int)(p / ((int64_t)A.t[2] * A.t[1]));
    const int oz = pz + A.off[0], oy = py + A.off[1], ox = px + A.off[2];
    const int nt = A.ntaps;
    const int64_t dvol = (int64_t)A.d[0] * A.d[1] * A.d[2];
    const float* highc = A.high + (int64_t)c * 7 * dvol;
    auto lo_i = [&](int o) { int v = o - nt + 1; return v <= 0 ? 0 : (v + 1) >> 1; };
    float acc = 0.0f;
    for (int iz = lo_i(oz); iz <= min(A.d[0] - 1, oz >> 1); ++iz) {
        const int tz = oz - 2 * iz;
        const float wz[2] = {A.lo[tz], A.hi[tz]};
        for (int iy = lo_i(oy); iy <= min(A.d[1] - 1, oy >> 1); ++iy) {
            const int ty = oy - 2 * iy;
            const float wy[2] = {A.lo[ty], A.hi[ty]};
            for (int ix = lo_i(ox); ix <= min(A.d[2] - 1, ox >> 1); ++ix) {
                const int tx = ox - 2 * ix;
                const float wx[2] = {A.lo[tx], A.hi[tx]};
                const int64_t b = ((int64_t)iz * A.d[1] + iy) * A.d[2] + ix;
                float v0 = A.low_cl ? A.low[b * A.C + c] : A.low[(int64_t)c * dvol + b];
                if (A.mlow) v0 *= A.mlow[b];
                acc = fmaf(v0, wz[0] * wy[0] * wx[0], acc);
#pragma unroll
                for (int k = 1; k < 8; ++k) {
                    float v = highc[(int64_t)(k - 1) * dvol + b];
                    if (A.mhigh) v *= A.mhigh[(int64_t)(k - 1) * dvol + b];
                    acc = fmaf(v, wz[(k >> 2) & 1] * wy[(k >> 1) & 1] * wx[k & 1], acc);
                }
            }
        }
    }
    A.out[idx] = acc;
}

static void launch_idwt_level(const LevelArgs& A, int64_t total, cudaStream_t st) {
    const unsigned blocks = (unsigned)((total + 127) / 128);
    if (A.ntaps == 2) (void)launch_pdl(idwt_level_kernel<2>, dim3(blocks), dim3(128), (size_t)(0), st, A);
    else if (A.ntaps == 4) (void)launch_pdl(idwt_level_kernel<4>, dim3(blocks), dim3(128), (size_t)(0), st, A);
    else if (A.ntaps == 6) (void)launch_pdl(idwt_level_kernel<6>, dim3(blocks), dim3(128), (size_t)(0), st, A);
    else (void)launch_pdl(idwt_level_generic_kernel, dim3(blocks), dim3(128), (size_t)(0), st, A);
}

// n_coeff == 1 (grid too small for a wavelet level): masked NCDHW -> channels-last copy
__global__ void copy_to_channels_last_kernel(const float* __restrict__ src, const float* __restrict__ mult,
                                             float* __restrict__ dst, float* __restrict__ also_zero, int C, int Cp,
                                             int64_t nvox) {
    LFGC_PDL_PROLOGUE();
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nvox * Cp) return;
    if (also_zero) also_zero[idx] = 0.0f;
    int c = (int)(idx % Cp);
    int64_t p = idx / Cp;
    float v = 0.0f;
    if (c < C) {
        v = src[(int64_t)c * nvox + p];
        if (mult) v *= mult[p];
    }
    dst[idx] = v;
}

struct LevelBwdArgs {
    const float* gout;     // d out, channels-last (t0,t1,t2,Cs)
    int Cs;
    const float* c_low;    // coefficient tensors of this level (needed for d mult); c_low only when low_is_coeff
    const float* c_high;
    const float* gm_low;   // gradient multipliers (null = 1)
    const float* gm_high;
    float* g_low;          // low_is_coeff ? grad of coeff[0] (C,d..) channels-first : scratch (d..,C) channels-last
    float* g_high;         // grad of coeff[l] (C,7,d..)
    float* gmult_low;      // d mult (d..) or null   (accumulated with atomics: zeroed by the host when !accumulate)
    float* gmult_high;     // (7,d..) or null
    int low_is_coeff;
    int accumulate;
    int C;
    int d[3], t[3], off[3];
    int ntaps;
    float lo[LFGC_MAX_TAPS], hi[LFGC_MAX_TAPS];
    // optional Adam epilogue on the detail bands of this level (lfgc_grid_step, split path): the thread that has just
    // formed the gradient of a coefficient updates it, so the gradient never makes a round trip through memory
    float* ap;             // coefficient tensor l inside the flat parameter buffer (null: no epilogue) ...
    float* am;             // ... and its Adam moments
    float* av;
    const float* lr;
    const int* step;       // optimiser steps taken so far (read only here; published by the grid-step kernel)
    AdamCoef coef;
    float w2x2;            // 2 * weight_l2 (SmallifyLoss weight term)
};

// One thread per (sub-band k, position b, channel c), channel fastest.  d mult[k][b] = sum_c coeff * g is reduced
// with warp shuffles when C divides the warp, then one atomic per (k, b).
template <int NT>
__global__ void idwt_level_bwd_kernel(LevelBwdArgs A) {
    LFGC_PDL_PROLOGUE();
    __shared__ float s_step_size, s_bc2_sqrt;
    if (A.ap) {
        if (threadIdx.x == 0) {
            float step_size, bc2_sqrt;
            adam_step_scalars(A.coef, *reinterpret_cast<const volatile int*>(A.step) + 1, *A.lr, step_size, bc2_sqrt);
            s_step_size = step_size;
            s_bc2_sqrt = bc2_sqrt;
        }
        __syncthreads();
    }
    const int64_t dvol = (int64_t)A.d[0] * A.d[1] * A.d[2];
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < 8 * dvol * A.C;
    const int c = live ? (int)(idx % A.C) : 0;
    const int64_t kb = live ? idx / A.C : 0;
    const int k = (int)(kb / dvol);
    const int64_t b = kb % dvol;
    const int ix = (int)(b % A.d[2]);
    const int iy = (int)((b / A.d[2]) % A.d[1]);
    const int iz = (int)(b / ((int64_t)A.d[2] * A.d[1]));
    const float* fz = ((k >> 2) & 1) ? A.hi : A.lo;
    const float* fy = ((k >> 1) & 1) ? A.hi : A.lo;
    const float* fx = (k & 1) ? A.hi : A.lo;
    // Adam epilogue: fetch the parameter and its moments up front, so that their round trip overlaps the gradient loads
    const bool adam = A.ap != nullptr && live && k > 0;
    const int64_t oa = ((int64_t)c * 7 + (k - 1)) * dvol + b;
    float a_p = 0.0f, a_m = 0.0f, a_v = 0.0f;
    if (adam) {
        a_p = A.ap[oa];
        a_m = A.am[oa];
        a_v = A.av[oa];
    }
    float g = 0.0f;
    if (live) {
        if (NT > 0) {
            // all NT^3 loads are independent: clamp the position, zero the weight when it falls outside the target
            int pp[3][NT > 0 ? NT : 1];
            float ww[3][NT > 0 ? NT : 1];
            const int i3[3] = {iz, iy, ix};
            const float* f3[3] = {fz, fy, fx};
#pragma unroll
            for (int ax = 0; ax < 3; ++ax)
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const int q = 2 * i3[ax] + t - A.off[ax];
                    const bool ok = q >= 0 && q < A.t[ax];
                    pp[ax][t] = ok ? q : 0;
                    ww[ax][t] = ok ? f3[ax][t] : 0.0f;
                }
#pragma unroll
            for (int tz = 0; tz < NT; ++tz)
#pragma unroll
                for (int ty = 0; ty < NT; ++ty) {
                    const float wzy = ww[0][tz] * ww[1][ty];
                    const int64_t rowp = ((int64_t)pp[0][tz] * A.t[1] + pp[1][ty]) * A.t[2];
#pragma unroll
                    for (int tx = 0; tx < NT; ++tx)
                        g = fmaf(A.gout[(rowp + pp[2][tx]) * A.Cs + c], wzy * ww[2][tx], g);
                }
        } else {
            const int nt = A.ntaps;
            for (int tz = 0; tz < nt; ++tz) {
                const int pz = 2 * iz + tz - A.off[0];
                if (pz < 0 || pz >= A.t[0]) continue;
                for (int ty = 0; ty < nt; ++ty) {
                    const int py = 2 * iy + ty - A.off[1];
                    if (py < 0 || py >= A.t[1]) continue;
                    const float wzy = fz[tz] * fy[ty];
                    for (int tx = 0; tx < nt; ++tx) {
                        const int px = 2 * ix + tx - A.off[2];
                        if (px < 0 || px >= A.t[2]) continue;
                        const int64_t p = ((int64_t)pz * A.t[1] + py) * A.t[2] + px;
                        g = fmaf(A.gout[p * A.Cs + c], wzy * fx[tx], g);
                    }
                }
            }
        }
    }
    float contrib = 0.0f;
    float* gm = nullptr;
    if (live) {
        if (k == 0) {
            if (A.low_is_coeff) {
                const int64_t o = (int64_t)c * dvol + b;
                if (A.gmult_low) { contrib = A.c_low[o] * g; gm = A.gmult_low + b; }
                const float v = A.gm_low ? g * A.gm_low[b] : g;
                A.g_low[o] = A.accumulate ? A.g_low[o] + v : v;
            } else {
                A.g_low[b * A.C + c] = g;
            }
        } else {
            const int64_t o = ((int64_t)c * 7 + (k - 1)) * dvol + b;
            if (A.gmult_high) { contrib = A.c_high[o] * g; gm = A.gmult_high + (int64_t)(k - 1) * dvol + b; }
            float v = A.gm_high ? g * A.gm_high[(int64_t)(k - 1) * dvol + b] : g;
            if (adam) {
                float pi = a_p, mi = a_m, vi = a_v;
                v = fmaf(A.w2x2, pi, v);
                adam_update(pi, v, mi, vi, A.coef, s_step_size, s_bc2_sqrt);
                A.ap[o] = pi;
                A.am[o] = mi;
                A.av[o] = vi;
                A.g_high[o] = v;
            } else {
                A.g_high[o] = A.accumulate ? A.g_high[o] + v : v;
            }
        }
    }
    if (A.gmult_low == nullptr && A.gmult_high == nullptr) return;  // uniform across the grid
    const int C = A.C;
    if (C <= 32 && (32 % C) == 0 && (blockDim.x % 32) == 0) {
        // a warp holds 32/C complete channel groups: segmented butterfly reduction
        for (int m = C >> 1; m >= 1; m >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, m);
        if (gm && c == 0) atomicAdd(gm, contrib);
    } else if (gm) {
        atomicAdd(gm, contrib);
    }
}

static void launch_idwt_level_bwd(const LevelBwdArgs& A, int64_t total, cudaStream_t st) {
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (A.ntaps == 2) (void)launch_pdl(idwt_level_bwd_kernel<2>, dim3(blocks), dim3(256), (size_t)(0), st, A);
    else if (A.ntaps == 4) (void)launch_pdl(idwt_level_bwd_kernel<4>, dim3(blocks), dim3(256), (size_t)(0), st, A);
    else (void)launch_pdl(idwt_level_bwd_kernel<0>, dim3(blocks), dim3(256), (size_t)(0), st, A);
}

// n_coeff == 1: adjoint of the masked transpose
__global__ void copy_from_channels_last_bwd_kernel(const float* __restrict__ gcl, const float* __restrict__ coeff,
                                                   const float* __restrict__ gmul, float* __restrict__ gcoeff,
                                                   float* __restrict__ gmult, int C, int Cp, int64_t nvox,
                                                   int accumulate) {
    LFGC_PDL_PROLOGUE();
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nvox) return;
    float msum = 0.0f;
    for (int c = 0; c < C; ++c) {
        const float g = gcl[p * Cp + c];
        const int64_t o = (int64_t)c * nvox + p;
        msum = fmaf(coeff[o], g, msum);
        const float v = gmul ? g * gmul[p] : g;
        gcoeff[o] = accumulate ? gcoeff[o] + v : v;
    }
    if (gmult) gmult[p] = accumulate ? gmult[p] + msum : msum;
}


// ---------------------------------------------------------------------------------------------------------------
// one analysis level (model construction only): _WaveletFilterNd.encode (Torch_Wavelet_Transform.py:75-89)
// zero padding (2L-3)//2 left, same (+1 if the extent is odd) right (:59-67), correlation with the flipped
// decomposition filters (:56), stride 2.
// ---------------------------------------------------------------------------------------------------------------
struct DwtArgs {
    const float* x;  // (C, d0, d1, d2)
    float* out;      // (C, 8, e0, e1, e2)
    int C;
    int d[3], e[3];
    int pad;
    int ntaps;
    float lo[LFGC_MAX_TAPS], hi[LFGC_MAX_TAPS];  // already flipped: filter_fwd taps
};

__global__ void dwt_level_kernel(DwtArgs A) {
    const int64_t evol = (int64_t)A.e[0] * A.e[1] * A.e[2];
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)A.C * 8 * evol) return;
    const int64_t b = idx % evol;
    const int k = (int)((idx / evol) % 8);
    const int c = (int)(idx / (8 * evol));
    const int ix = (int)(b % A.e[2]);
    const int iy = (int)((b / A.e[2]) % A.e[1]);
    const int iz = (int)(b / ((int64_t)A.e[2] * A.e[1]));
    const float* fz = ((k >> 2) & 1) ? A.hi : A.lo;
    const float* fy = ((k >> 1) & 1) ? A.hi : A.lo;
    const float* fx = (k & 1) ? A.hi : A.lo;
    const float* xc = A.x + (int64_t)c * A.d[0] * A.d[1] * A.d[2];
    float acc = 0.0f;
    for (int tz = 0; tz < A.ntaps; ++tz) {
        const int z = 2 * iz + tz - A.pad;
        if (z < 0 || z >= A.d[0]) continue;
        for (int ty = 0; ty < A.ntaps; ++ty) {
            const int y = 2 * iy + ty - A.pad;
            if (y < 0 || y >= A.d[1]) continue;
            const float wzy = fz[tz] * fy[ty];
            for (int tx = 0; tx < A.ntaps; ++tx) {
                const int x = 2 * ix + tx - A.pad;
                if (x < 0 || x >= A.d[2]) continue;
                acc = fmaf(xc[((int64_t)z * A.d[1] + y) * A.d[2] + x], wzy * fx[tx], acc);
            }
        }
    }
    A.out[idx] = acc;
}

static int check_desc(const lfgc_wavelet_desc* w) {
    if (!w) return fail(LFGC_E_INVALID, "wavelet desc is null");
    if (w->n_coeff < 1 || w->n_coeff > LFGC_MAX_LEVELS) return fail(LFGC_E_INVALID, "n_coeff=%d out of range", w->n_coeff);
    if (w->C < 1) return fail(LFGC_E_INVALID, "C=%d", w->C);
    if (w->n_coeff > 1 && (w->n_taps < 2 || w->n_taps > LFGC_MAX_TAPS || (w->n_taps & 1)))
        return fail(LFGC_E_UNSUPPORTED, "filter length %d unsupported (even, <= %d)", w->n_taps, LFGC_MAX_TAPS);
    for (int l = 1; l < w->n_coeff; ++l)
        for (int a = 0; a < 3; ++a) {
            const int full = 2 * w->dims[l][a] + w->n_taps - 2;
            if (w->target[l][a] < 1 || w->target[l][a] > full)
                return fail(LFGC_E_INVALID, "level %d: target %d does not fit synthesis size %d", l, w->target[l][a], full);
            if (l >= 2 && w->dims[l][a] != w->target[l - 1][a])
                return fail(LFGC_E_INVALID, "level %d: dims must equal the previous level's target", l);
            if (l == 1 && w->dims[1][a] != w->dims[0][a])
                return fail(LFGC_E_INVALID, "level 1 dims must equal the low-pass dims");
        }
    return LFGC_OK;
}

// wavelet_sep.cu: separable two-kernels-per-level path for large mask-free pyramids
size_t wavelet_sep_scratch_elems(const lfgc_wavelet_desc* w);
bool wavelet_sep_preferred(const lfgc_wavelet_desc* w);
int wavelet_sep_fwd(const lfgc_wavelet_desc* w, const float* const* coeff, float* scratch, float* grid_cl, int Cp,
                    float* also_zero, cudaStream_t st);
int wavelet_sep_bwd(const lfgc_wavelet_desc* w, const float* grad_grid_cl, int Cp, float* scratch, float* const* grad_coeff,
                    cudaStream_t st);

static size_t intermediate_elems(const lfgc_wavelet_desc* w) {
    size_t mx = 0;
    for (int l = 1; l < w->n_coeff - 1; ++l) {
        size_t e = (size_t)w->C * w->target[l][0] * w->target[l][1] * w->target[l][2];
        if (e > mx) mx = e;
    }
    return mx;
}

static void fill_level_geometry(const lfgc_wavelet_desc* w, int l, int (&d)[3], int (&t)[3], int (&off)[3], float (&lo)[LFGC_MAX_TAPS],
                                float (&hi)[LFGC_MAX_TAPS]) {
    for (int a = 0; a < 3; ++a) {
        d[a] = w->dims[l][a];
        t[a] = w->target[l][a];
        off[a] = (2 * d[a] + w->n_taps - 2 - t[a]) / 2;  // floor(delta / 2), Torch_Wavelet_Transform.py:71
    }
    for (int i = 0; i < LFGC_MAX_TAPS; ++i) {
        lo[i] = w->rec_lo[i];
        hi[i] = w->rec_hi[i];
    }
}

// Finest-level adjoint with the Adam update of that level's detail bands in its epilogue (mask-free).  grad_grid_cl:
// (t0,t1,t2,Cp) channels-last; low_grad_out: gradient of the level's low-pass input, channels-last (d0,d1,d2,C).
int wavelet_finest_bwd_adam(const lfgc_wavelet_desc* w, const float* grad_grid_cl, int Cp, float* low_grad_out, float* p_l,
                            float* g_l, float* m_l, float* v_l, const float* lr, const int* step, const AdamCoef& coef,
                            float w2x2, cudaStream_t st) {
    const int l = w->n_coeff - 1;
    LevelBwdArgs A = {};
    A.gout = grad_grid_cl;
    A.Cs = Cp;
    A.low_is_coeff = 0;
    A.c_low = nullptr;
    A.c_high = p_l;
    A.g_low = low_grad_out;
    A.g_high = g_l;
    A.accumulate = 0;
    A.C = w->C;
    A.ntaps = w->n_taps;
    fill_level_geometry(w, l, A.d, A.t, A.off, A.lo, A.hi);
    A.ap = p_l;
    A.am = m_l;
    A.av = v_l;
    A.lr = lr;
    A.step = step;
    A.coef = coef;
    A.w2x2 = w2x2;
    const int64_t total = 8 * (int64_t)A.d[0] * A.d[1] * A.d[2] * A.C;
    launch_idwt_level_bwd(A, total, st);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

// Finest-level synthesis from a channels-last low-pass input (d0,d1,d2,C) and the level's detail bands.
int wavelet_finest_fwd(const lfgc_wavelet_desc* w, const float* low_cl, const float* coeff_l, float* grid_cl, int Cp,
                       float* also_zero, cudaStream_t st) {
    const int l = w->n_coeff - 1;
    LevelArgs A = {};
    A.low = low_cl;
    A.high = coeff_l;
    A.out = grid_cl;
    A.also_zero = also_zero;
    A.low_cl = 1;
    A.C = w->C;
    A.Cs = Cp;
    A.ntaps = w->n_taps;
    fill_level_geometry(w, l, A.d, A.t, A.off, A.lo, A.hi);
    const int64_t total = (int64_t)A.t[0] * A.t[1] * A.t[2] * A.Cs;
    launch_idwt_level(A, total, st);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

}  // namespace lfgc

using namespace lfgc;

extern "C" int lfgc_mask_multiplier(int mode, int64_t n, const float* p0, const float* p1, const float* noise,
                                    float threshold, float* mult_out, float* aux_out, void* stream) {
    if (n < 0 || !mult_out) return fail(LFGC_E_INVALID, "mask_multiplier: bad arguments");
    if (mode < LFGC_MASK_IDENTITY || mode > LFGC_MASK_BERNOULLI) return fail(LFGC_E_INVALID, "mask mode %d", mode);
    if (mode != LFGC_MASK_IDENTITY && !p0) return fail(LFGC_E_INVALID, "mask_multiplier: p0 is null");
    if (mode == LFGC_MASK_VARIATIONAL && (!p1 || !noise)) return fail(LFGC_E_INVALID, "variational mask needs p1 and noise");
    if (mode == LFGC_MASK_BERNOULLI && !noise) return fail(LFGC_E_INVALID, "bernoulli mask needs noise");
    if (n == 0) return LFGC_OK;
    (void)launch_pdl(mask_multiplier_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, mode, n, p0, p1, noise,
                                                                                         threshold, mult_out, aux_out);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

extern "C" int lfgc_mask_param_grad(int mode, int64_t n, const float* p0, const float* p1, const float* noise,
                                    const float* gmult, float* g0, float* g1, int accumulate, void* stream) {
    if (n < 0 || !gmult) return fail(LFGC_E_INVALID, "mask_param_grad: bad arguments");
    if (mode == LFGC_MASK_IDENTITY || mode == LFGC_MASK_BERNOULLI || n == 0) return LFGC_OK;
    if (mode < 0 || mode > LFGC_MASK_BERNOULLI) return fail(LFGC_E_INVALID, "mask mode %d", mode);
    if (!g0 || !p0) return fail(LFGC_E_INVALID, "mask_param_grad: p0/g0 null");
    if (mode == LFGC_MASK_VARIATIONAL && (!p1 || !g1 || !noise)) return fail(LFGC_E_INVALID, "variational grad needs p1,g1,noise");
    (void)launch_pdl(mask_param_grad_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, mode, n, p0, p1, noise, gmult,
                                                                                         g0, g1, accumulate);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

extern "C" int lfgc_smallify_ema(const float* betas, float* ema, float* emavar, int64_t n, float momentum,
                                 void* stream) {
    if (!betas || !ema || !emavar || n < 0) return fail(LFGC_E_INVALID, "smallify_ema: bad arguments");
    if (n == 0) return LFGC_OK;
    (void)launch_pdl(smallify_ema_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, betas, ema, emavar, n, momentum);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

extern "C" size_t lfgc_decode_scratch_bytes(const lfgc_wavelet_desc* w) {
    if (!w || check_desc(w) != LFGC_OK) return 0;
    size_t elems = 2 * intermediate_elems(w);
    if (wavelet_sep_scratch_elems(w) > elems) elems = wavelet_sep_scratch_elems(w);   // whichever path is taken at run time
    return elems * sizeof(float) + 16;
}

extern "C" int lfgc_decode_fwd(const lfgc_wavelet_desc* w, const float* const* coeff, const float* const* mult,
                               float* scratch, float* grid_cl, int Cp, float* also_zero, void* stream) {
    int rc = check_desc(w);
    if (rc) return rc;
    if (!coeff || !grid_cl) return fail(LFGC_E_INVALID, "decode_fwd: null pointer");
    if (Cp < w->C || (Cp & 3)) return fail(LFGC_E_INVALID, "Cp=%d must be a multiple of 4 and >= C=%d", Cp, w->C);
    cudaStream_t st = (cudaStream_t)stream;
    if (w->n_coeff == 1) {
        const int64_t nvox = (int64_t)w->dims[0][0] * w->dims[0][1] * w->dims[0][2];
        (void)launch_pdl(copy_to_channels_last_kernel, dim3((unsigned)((nvox * Cp + 255) / 256)), dim3(256), (size_t)(0), st, coeff[0], mult ? mult[0] : nullptr,
                                                                                        grid_cl, also_zero, w->C, Cp, nvox);
        LFGC_LAUNCH_OK();
        return LFGC_OK;
    }
    {   // large mask-free pyramids: separable passes (wavelet_sep.cu)
        bool masked = false;
        for (int l = 0; mult && l < w->n_coeff; ++l) masked = masked || mult[l] != nullptr;
        if (!masked && wavelet_sep_preferred(w)) {
            if (!scratch) return fail(LFGC_E_WORKSPACE, "decode_fwd: scratch required");
            return wavelet_sep_fwd(w, coeff, scratch, grid_cl, Cp, also_zero, st);
        }
    }
    const size_t inter = intermediate_elems(w);
    if (inter && !scratch) return fail(LFGC_E_WORKSPACE, "decode_fwd: scratch required");
    float* buf[2] = {scratch, scratch ? scratch + inter : nullptr};
    const float* low = coeff[0];
    for (int l = 1; l < w->n_coeff; ++l) {
        LevelArgs A;
        A.low = low;
        A.high = coeff[l];
        A.mlow = (l == 1 && mult) ? mult[0] : nullptr;
        A.mhigh = mult ? mult[l] : nullptr;
        const bool last = (l == w->n_coeff - 1);
        A.out = last ? grid_cl : buf[l & 1];
        A.also_zero = last ? also_zero : nullptr;
        A.low_cl = (l == 1) ? 0 : 1;
        A.C = w->C;
        A.Cs = last ? Cp : w->C;
        A.ntaps = w->n_taps;
        for (int a = 0; a < 3; ++a) {
            A.d[a] = w->dims[l][a];
            A.t[a] = w->target[l][a];
            A.off[a] = (2 * A.d[a] + A.ntaps - 2 - A.t[a]) / 2;  // floor(delta / 2), Torch_Wavelet_Transform.py:71
        }
        for (int i = 0; i < LFGC_MAX_TAPS; ++i) {
            A.lo[i] = w->rec_lo[i];
            A.hi[i] = w->rec_hi[i];
        }
        const int64_t total = (int64_t)A.t[0] * A.t[1] * A.t[2] * A.Cs;
        launch_idwt_level(A, total, st);
        LFGC_LAUNCH_OK();
        low = A.out;
    }
    return LFGC_OK;
}

extern "C" int lfgc_decode_bwd(const lfgc_wavelet_desc* w, const float* grad_grid_cl, int Cp,
                               const float* const* coeff, const float* const* gmul, float* scratch,
                               float* const* grad_coeff, float* const* grad_mult, int accumulate, void* stream) {
    int rc = check_desc(w);
    if (rc) return rc;
    if (!coeff || !grad_grid_cl || !grad_coeff) return fail(LFGC_E_INVALID, "decode_bwd: null pointer");
    if (Cp < w->C || (Cp & 3)) return fail(LFGC_E_INVALID, "Cp=%d must be a multiple of 4 and >= C=%d", Cp, w->C);
    cudaStream_t st = (cudaStream_t)stream;
    if (w->n_coeff == 1) {
        const int64_t nvox = (int64_t)w->dims[0][0] * w->dims[0][1] * w->dims[0][2];
        (void)launch_pdl(copy_from_channels_last_bwd_kernel, dim3((unsigned)((nvox + 127) / 128)), dim3(128), (size_t)(0), st, 
            grad_grid_cl, coeff[0], gmul ? gmul[0] : nullptr, grad_coeff[0], grad_mult ? grad_mult[0] : nullptr, w->C,
            Cp, nvox, accumulate);
        LFGC_LAUNCH_OK();
        return LFGC_OK;
    }
    {   // large mask-free pyramids: separable passes (wavelet_sep.cu)
        bool masked = false;
        for (int l = 0; l < w->n_coeff; ++l)
            masked = masked || (gmul && gmul[l] != nullptr) || (grad_mult && grad_mult[l] != nullptr);
        if (!masked && !accumulate && wavelet_sep_preferred(w)) {
            if (!scratch) return fail(LFGC_E_WORKSPACE, "decode_bwd: scratch required");
            for (int l = 0; l < w->n_coeff; ++l)
                if (!grad_coeff[l]) return fail(LFGC_E_INVALID, "decode_bwd: grad_coeff[%d] null", l);
            return wavelet_sep_bwd(w, grad_grid_cl, Cp, scratch, grad_coeff, st);
        }
    }
    const size_t inter = intermediate_elems(w);
    if (inter && !scratch) return fail(LFGC_E_WORKSPACE, "decode_bwd: scratch required");
    float* buf[2] = {scratch, scratch ? scratch + inter : nullptr};
    const float* gout = grad_grid_cl;
    for (int l = w->n_coeff - 1; l >= 1; --l) {
        LevelBwdArgs A = {};
        A.gout = gout;
        A.Cs = (l == w->n_coeff - 1) ? Cp : w->C;
        A.low_is_coeff = (l == 1) ? 1 : 0;
        A.c_low = coeff[0];
        A.c_high = coeff[l];
        A.gm_low = (l == 1 && gmul) ? gmul[0] : nullptr;
        A.gm_high = gmul ? gmul[l] : nullptr;
        A.g_low = (l == 1) ? grad_coeff[0] : buf[l & 1];
        A.g_high = grad_coeff[l];
        A.gmult_low = (l == 1 && grad_mult) ? grad_mult[0] : nullptr;
        A.gmult_high = grad_mult ? grad_mult[l] : nullptr;
        A.accumulate = accumulate == 1 ? 1 : 0;   // 2: grad_mult pre-cleared by the caller, grad_coeff overwritten
        A.C = w->C;
        A.ntaps = w->n_taps;
        for (int a = 0; a < 3; ++a) {
            A.d[a] = w->dims[l][a];
            A.t[a] = w->target[l][a];
            A.off[a] = (2 * A.d[a] + A.ntaps - 2 - A.t[a]) / 2;
        }
        for (int i = 0; i < LFGC_MAX_TAPS; ++i) {
            A.lo[i] = w->rec_lo[i];
            A.hi[i] = w->rec_hi[i];
        }
        if (!A.g_low || !A.g_high) return fail(LFGC_E_INVALID, "decode_bwd: grad_coeff[%d] null", l);
        const int64_t dvol = (int64_t)A.d[0] * A.d[1] * A.d[2];
        if (!accumulate) {  // d mult is accumulated with atomics
            if (A.gmult_low) LFGC_CUDA_OK(cudaMemsetAsync(A.gmult_low, 0, dvol * sizeof(float), st));
            if (A.gmult_high) LFGC_CUDA_OK(cudaMemsetAsync(A.gmult_high, 0, 7 * dvol * sizeof(float), st));
        }
        const int64_t total = 8 * dvol * A.C;
        launch_idwt_level_bwd(A, total, st);
        LFGC_LAUNCH_OK();
        gout = A.g_low;
    }
    return LFGC_OK;
}

extern "C" int lfgc_dwt_level(const float* x, int C, const int32_t d[3], int n_taps, const float* dec_lo,
                              const float* dec_hi, float* out, int32_t e_out[3], void* stream) {
    if (!d || !dec_lo || !dec_hi || !e_out) return fail(LFGC_E_INVALID, "dwt_level: null pointer");
    if (n_taps < 2 || n_taps > LFGC_MAX_TAPS || (n_taps & 1)) return fail(LFGC_E_UNSUPPORTED, "filter length %d", n_taps);
    DwtArgs A;
    A.x = x;
    A.out = out;
    A.C = C;
    A.ntaps = n_taps;
    A.pad = (2 * n_taps - 3) / 2;
    for (int a = 0; a < 3; ++a) {
        if (d[a] < 1) return fail(LFGC_E_INVALID, "dwt_level: bad extent");
        A.d[a] = d[a];
        const int padded = d[a] + 2 * A.pad + (d[a] & 1);
        A.e[a] = (padded - n_taps) / 2 + 1;
        e_out[a] = A.e[a];
    }
    if (!x || !out) return LFGC_OK;  // size query only
    for (int i = 0; i < LFGC_MAX_TAPS; ++i) {
        A.lo[i] = i < n_taps ? dec_lo[n_taps - 1 - i] : 0.0f;  // flipped (:56)
        A.hi[i] = i < n_taps ? dec_hi[n_taps - 1 - i] : 0.0f;
    }
    const int64_t total = (int64_t)C * 8 * A.e[0] * A.e[1] * A.e[2];
    dwt_level_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(A);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}
