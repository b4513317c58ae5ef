// "Glue" kernel: everything of one optimiser step that is NOT per-sample, for mask-free models, in ONE cooperative
// launch instead of five dependent ones:
//
//     synthesis adjoint, finest -> coarsest level     (autograd of Feature_Grid_Model.decode_volume,
//                                                      model/Feature_Grid_Model.py:102-108, training/training.py:137)
//     Adam over the flat parameter buffer             (torch.optim.Adam, training/training.py:199,232)
//     synthesis, coarsest -> finest level, of the UPDATED coefficients: the grid the next step's fused training
//     kernel gathers from, and the zeroed gradient accumulator it scatters into
//
// The phases are separated by grid-wide barriers (cooperative groups) instead of kernel boundaries: with the
// 0.45 MB working set of the shipped configurations each phase is a few microseconds of work, so the dependent-launch
// gaps of the five separate kernels (~22 us per step inside the CUDA graph, profiles/README.md) dominate them.
// Per-element arithmetic is the one of wavelet.cu (idwt_level_kernel / idwt_level_bwd_kernel without mask
// multipliers) and optim.cu (adam_kernel): same summation order, same results.
//
// Opt-in (FastTrainer, LFGC_GLUE=1) until it has been measured on a B200; the separate kernels stay the default.
#include "lfgc_common.cuh"

#include <cooperative_groups.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace lfgc {
namespace glue {

struct Args {
    int n_coeff, C, Cp, ntaps;
    float lo[LFGC_MAX_TAPS], hi[LFGC_MAX_TAPS];
    int d[LFGC_MAX_LEVELS][3], t[LFGC_MAX_LEVELS][3], off[LFGC_MAX_LEVELS][3];
    float* coeff[LFGC_MAX_LEVELS];    // coefficient tensors (reference layout, inside the flat parameter buffer)
    float* gcoeff[LFGC_MAX_LEVELS];   // their gradients
    float* buf[2];                    // channels-last intermediates (decode scratch)
    const float* grad_grid;           // d loss / d grid, channels-last (G0,G1,G2,Cp)
    float* grid_cl;                   // decoded grid, channels-last
    float* also_zero;                 // gradient accumulator cleared together with the last synthesis level
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
    const float* lr;
    int* step;
    AdamCoef c;
    int phases;                       // 1 = adjoint, 2 = Adam, 4 = synthesis
};

// one output element of synthesis level l (wavelet.cu idwt_level_kernel<NT>, no multipliers)
template <int NT>
__host__ __device__ __forceinline__ void synth_elem(const Args& A, int l, long long idx) {
    constexpr int NP = NT / 2;
    const bool last = l == A.n_coeff - 1;
    const int Cs = last ? A.Cp : A.C;
    const int* t = A.t[l];
    const int* d = A.d[l];
    const long long nvox = (long long)t[0] * t[1] * t[2];
    if (idx >= nvox * Cs) return;
    float* out = last ? A.grid_cl : A.buf[l & 1];
    const int c = (int)(idx % Cs);
    const long long p = idx / Cs;
    if (last && A.also_zero) A.also_zero[idx] = 0.0f;
    if (c >= A.C) {
        out[idx] = 0.0f;  // zero pad channels
        return;
    }
    const int px = (int)(p % t[2]);
    const int py = (int)((p / t[2]) % t[1]);
    const int pz = (int)(p / ((long long)t[2] * t[1]));
    const int o[3] = {pz + A.off[l][0], py + A.off[l][1], px + A.off[l][2]};
    const long long dvol = (long long)d[0] * d[1] * d[2];
    const float* low = l == 1 ? A.coeff[0] : A.buf[(l - 1) & 1];
    const bool low_cl = l > 1;
    const float* highc = A.coeff[l] + (long long)c * 7 * dvol;

    int ii[3][NP];
    float wl[3][NP], wh[3][NP];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax)
#pragma unroll
        for (int a = 0; a < NP; ++a) {
            const int i = (o[ax] >> 1) - a;
            const int tt = (o[ax] & 1) + 2 * a;
            const bool ok = i >= 0 && i < d[ax];
            ii[ax][a] = ok ? i : 0;
            wl[ax][a] = ok ? A.lo[tt] : 0.0f;
            wh[ax][a] = ok ? A.hi[tt] : 0.0f;
        }
    float acc = 0.0f;
#pragma unroll
    for (int az = 0; az < NP; ++az)
#pragma unroll
        for (int ay = 0; ay < NP; ++ay)
#pragma unroll
            for (int axx = 0; axx < NP; ++axx) {
                const long long b = ((long long)ii[0][az] * d[1] + ii[1][ay]) * d[2] + ii[2][axx];
                float vv[8];
                vv[0] = low_cl ? low[b * A.C + c] : low[(long long)c * dvol + b];
#pragma unroll
                for (int k = 1; k < 8; ++k) vv[k] = highc[(long long)(k - 1) * dvol + b];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float wz = (k & 4) ? wh[0][az] : wl[0][az];
                    const float wy = (k & 2) ? wh[1][ay] : wl[1][ay];
                    const float wx = (k & 1) ? wh[2][axx] : wl[2][axx];
                    acc = fmaf(vv[k], wz * wy * wx, acc);
                }
            }
    out[idx] = acc;
}

// one (sub-band, position, channel) element of the adjoint of level l (wavelet.cu idwt_level_bwd_kernel<NT>, no
// multipliers, overwrite semantics)
template <int NT>
__host__ __device__ __forceinline__ void adjoint_elem(const Args& A, int l, long long idx) {
    const int* t = A.t[l];
    const int* d = A.d[l];
    const long long dvol = (long long)d[0] * d[1] * d[2];
    if (idx >= 8 * dvol * A.C) return;
    const bool finest = l == A.n_coeff - 1;
    const float* gout = finest ? A.grad_grid : A.buf[(l + 1) & 1];
    const int Cs = finest ? A.Cp : A.C;
    const int c = (int)(idx % A.C);
    const long long kb = idx / A.C;
    const int k = (int)(kb / dvol);
    const long long b = kb % dvol;
    const int ix = (int)(b % d[2]);
    const int iy = (int)((b / d[2]) % d[1]);
    const int iz = (int)(b / ((long long)d[2] * d[1]));
    const float* fz = ((k >> 2) & 1) ? A.hi : A.lo;
    const float* fy = ((k >> 1) & 1) ? A.hi : A.lo;
    const float* fx = (k & 1) ? A.hi : A.lo;
    int pp[3][NT];
    float ww[3][NT];
    const int i3[3] = {iz, iy, ix};
    const float* f3[3] = {fz, fy, fx};
#pragma unroll
    for (int ax = 0; ax < 3; ++ax)
#pragma unroll
        for (int tt = 0; tt < NT; ++tt) {
            const int q = 2 * i3[ax] + tt - A.off[l][ax];
            const bool ok = q >= 0 && q < t[ax];
            pp[ax][tt] = ok ? q : 0;
            ww[ax][tt] = ok ? f3[ax][tt] : 0.0f;
        }
    float g = 0.0f;
#pragma unroll
    for (int tz = 0; tz < NT; ++tz)
#pragma unroll
        for (int ty = 0; ty < NT; ++ty) {
            const float wzy = ww[0][tz] * ww[1][ty];
            const long long rowp = ((long long)pp[0][tz] * t[1] + pp[1][ty]) * t[2];
#pragma unroll
            for (int tx = 0; tx < NT; ++tx) g = fmaf(gout[(rowp + pp[2][tx]) * Cs + c], wzy * ww[2][tx], g);
        }
    if (k == 0) {
        if (l == 1) A.gcoeff[0][(long long)c * dvol + b] = g;          // gradient of the coarsest low-pass tensor
        else A.buf[l & 1][b * A.C + c] = g;                             // d(low-pass input), channels-last
    } else {
        A.gcoeff[l][((long long)c * 7 + (k - 1)) * dvol + b] = g;
    }
}

// torch.optim.Adam for one element: the function optim.cu's adam_kernel uses (lfgc_common.cuh adam_update)
__host__ __device__ __forceinline__ void adam_elem(const Args& A, long long i, float step_size, float bc2_sqrt) {
    float pi = A.p[i], mi = A.m[i], vi = A.v[i];
    adam_update(pi, A.g[i], mi, vi, A.c, step_size, bc2_sqrt);
    A.p[i] = pi;
    A.m[i] = mi;
    A.v[i] = vi;
}

template <int NT>
__global__ void __launch_bounds__(256) step_glue_kernel(const __grid_constant__ Args A) {
    cg::grid_group grid = cg::this_grid();
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsize = (long long)gridDim.x * blockDim.x;

    if (A.phases & 1) {
        if (A.n_coeff == 1) {
            // no wavelet level: adjoint of the NCDHW -> channels-last copy
            const long long nvox = (long long)A.d[0][0] * A.d[0][1] * A.d[0][2];
            for (long long i = gtid; i < nvox * A.C; i += gsize) {
                const int c = (int)(i % A.C);
                const long long p = i / A.C;
                A.gcoeff[0][(long long)c * nvox + p] = A.grad_grid[p * A.Cp + c];
            }
        } else {
            for (int l = A.n_coeff - 1; l >= 1; --l) {
                const long long total = 8ll * A.d[l][0] * A.d[l][1] * A.d[l][2] * A.C;
                for (long long i = gtid; i < total; i += gsize) adjoint_elem<NT>(A, l, i);
                if (l > 1) grid.sync();   // the next (coarser) level reads this level's low-pass gradient
            }
        }
        if (A.phases & 6) grid.sync();
    }

    if (A.phases & 2) {
        // torch.optim.Adam; every thread reads the step count before anybody publishes the new one (after the barrier)
        const int step = *reinterpret_cast<volatile int*>(A.step) + 1;
        float step_size, bc2_sqrt;
        adam_step_scalars(A.c, step, *A.lr, step_size, bc2_sqrt);
        for (long long i = gtid; i < A.n; i += gsize) adam_elem(A, i, step_size, bc2_sqrt);
        grid.sync();
        if (gtid == 0) *A.step = step;
    }

    if (A.phases & 4) {
        if (A.n_coeff == 1) {
            const long long nvox = (long long)A.d[0][0] * A.d[0][1] * A.d[0][2];
            for (long long i = gtid; i < nvox * A.Cp; i += gsize) {
                const int c = (int)(i % A.Cp);
                const long long p = i / A.Cp;
                if (A.also_zero) A.also_zero[i] = 0.0f;
                A.grid_cl[i] = c < A.C ? A.coeff[0][(long long)c * nvox + p] : 0.0f;
            }
        } else {
            for (int l = 1; l < A.n_coeff; ++l) {
                const bool last = l == A.n_coeff - 1;
                const long long total = (long long)A.t[l][0] * A.t[l][1] * A.t[l][2] * (last ? A.Cp : A.C);
                for (long long i = gtid; i < total; i += gsize) synth_elem<NT>(A, l, i);
                if (!last) grid.sync();
            }
        }
    }
}

template <int NT>
static int launch(const Args& A, cudaStream_t st) {
    auto kern = step_glue_kernel<NT>;
    int occ = 0;
    LFGC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
    if (occ < 1) return fail(LFGC_E_UNSUPPORTED, "step_glue: kernel does not fit on an SM");
    int cap = 2;   // CTAs per SM: more CTAs shorten the phases, fewer shorten the grid barriers (tuning: LFGC_GLUE_OCC)
    if (const char* e = getenv("LFGC_GLUE_OCC")) { const int v = atoi(e); if (v >= 1 && v <= 8) cap = v; }
    if (occ > cap) occ = cap;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(sm_count() * occ));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    LFGC_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, A));
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

}  // namespace glue
}  // namespace lfgc

using namespace lfgc;

static int glue_fill_args(glue::Args& A, const lfgc_wavelet_desc* w, int Cp, float* const* coeff, float* const* grad_coeff,
                          float* scratch, const float* grad_grid_cl, float* grid_cl, float* also_zero, float* p,
                          const float* g, float* m, float* v, int64_t n, const float* lr, int32_t* step_count,
                          double beta1, double beta2, double eps, double grad_scale, int phases) {

    if (!w || !coeff || !grad_coeff) return fail(LFGC_E_INVALID, "step_glue: null descriptor / pointer table");
    if (w->n_coeff < 1 || w->n_coeff > LFGC_MAX_LEVELS || w->C < 1) return fail(LFGC_E_INVALID, "step_glue: bad descriptor");
    if (w->n_coeff > 1 && w->n_taps != 2 && w->n_taps != 4)
        return fail(LFGC_E_UNSUPPORTED, "step_glue: filter length %d (the fused step covers 2 and 4 taps)", w->n_taps);
    if (Cp < w->C || (Cp & 3)) return fail(LFGC_E_INVALID, "Cp=%d must be a multiple of 4 and >= C=%d", Cp, w->C);
    if ((phases & 1) && !grad_grid_cl) return fail(LFGC_E_INVALID, "step_glue: grad_grid_cl is null");
    if ((phases & 2) && (!p || !g || !m || !v || !lr || !step_count || n < 0)) return fail(LFGC_E_INVALID, "step_glue: bad Adam arguments");
    if ((phases & 4) && !grid_cl) return fail(LFGC_E_INVALID, "step_glue: grid_cl is null");
    A.n_coeff = w->n_coeff;
    A.C = w->C;
    A.Cp = Cp;
    A.ntaps = w->n_taps;
    for (int i = 0; i < LFGC_MAX_TAPS; ++i) {
        A.lo[i] = w->rec_lo[i];
        A.hi[i] = w->rec_hi[i];
    }
    size_t inter = 0;   // largest channels-last intermediate (levels 1 .. n_coeff-2)
    for (int l = 0; l < LFGC_MAX_LEVELS; ++l) {
        for (int a = 0; a < 3; ++a) {
            A.d[l][a] = l < w->n_coeff ? w->dims[l][a] : 0;
            A.t[l][a] = l < w->n_coeff ? w->target[l][a] : 0;
            A.off[l][a] = (l >= 1 && l < w->n_coeff) ? (2 * A.d[l][a] + A.ntaps - 2 - A.t[l][a]) / 2 : 0;
        }
        A.coeff[l] = l < w->n_coeff ? coeff[l] : nullptr;
        A.gcoeff[l] = l < w->n_coeff ? grad_coeff[l] : nullptr;
        if (l < w->n_coeff && (!A.coeff[l] || ((phases & 1) && !A.gcoeff[l])))
            return fail(LFGC_E_INVALID, "step_glue: coefficient / gradient pointer %d is null", l);
        if (l >= 1 && l < w->n_coeff - 1) {
            const size_t e = (size_t)A.t[l][0] * A.t[l][1] * A.t[l][2] * w->C;
            if (e > inter) inter = e;
        }
    }
    if (inter && !scratch) return fail(LFGC_E_WORKSPACE, "step_glue: scratch required");
    A.buf[0] = scratch;
    A.buf[1] = scratch ? scratch + inter : nullptr;
    A.grad_grid = grad_grid_cl;
    A.grid_cl = grid_cl;
    A.also_zero = also_zero;
    A.p = p;
    A.g = g;
    A.m = m;
    A.v = v;
    A.n = n;
    A.lr = lr;
    A.step = step_count;
    A.c = make_adam_coef(beta1, beta2, eps, grad_scale);
    A.phases = phases & 7;
    return LFGC_OK;
}

extern "C" int lfgc_step_glue(const lfgc_wavelet_desc* w, int Cp, float* const* coeff, float* const* grad_coeff,
                              float* scratch, const float* grad_grid_cl, float* grid_cl, float* also_zero, float* p,
                              const float* g, float* m, float* v, int64_t n, const float* lr, int32_t* step_count,
                              double beta1, double beta2, double eps, double grad_scale, int phases, void* stream) {
    if ((phases & 7) == 0) return LFGC_OK;
    glue::Args A;
    const int rc = glue_fill_args(A, w, Cp, coeff, grad_coeff, scratch, grad_grid_cl, grid_cl, also_zero, p, g, m, v, n, lr,
                                  step_count, beta1, beta2, eps, grad_scale, phases);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return w->n_taps == 2 || w->n_coeff == 1 ? glue::launch<2>(A, st) : glue::launch<4>(A, st);
}

#ifdef LFGC_GLUE_HOST_TEST
// Test hook (only in builds made by tests/test_glue_host.py, never in liblfgc.so): the SAME per-element functions and the
// same phase order run sequentially on HOST memory, so the index arithmetic of the glue kernel can be checked against
// the numpy oracle without a GPU.  All pointers are host pointers here.
template <int NT>
static void glue_run_host(const glue::Args& A) {
    using namespace glue;
    if (A.phases & 1) {
        if (A.n_coeff == 1) {
            const long long nvox = (long long)A.d[0][0] * A.d[0][1] * A.d[0][2];
            for (long long i = 0; i < nvox * A.C; ++i) A.gcoeff[0][(i % A.C) * nvox + i / A.C] = A.grad_grid[(i / A.C) * A.Cp + i % A.C];
        } else {
            for (int l = A.n_coeff - 1; l >= 1; --l) {
                const long long total = 8ll * A.d[l][0] * A.d[l][1] * A.d[l][2] * A.C;
                for (long long i = 0; i < total; ++i) adjoint_elem<NT>(A, l, i);
            }
        }
    }
    if (A.phases & 2) {
        const int step = *A.step + 1;
        float step_size, bc2_sqrt;
        adam_step_scalars(A.c, step, *A.lr, step_size, bc2_sqrt);
        for (long long i = 0; i < A.n; ++i) adam_elem(A, i, step_size, bc2_sqrt);
        *A.step = step;
    }
    if (A.phases & 4) {
        if (A.n_coeff == 1) {
            const long long nvox = (long long)A.d[0][0] * A.d[0][1] * A.d[0][2];
            for (long long i = 0; i < nvox * A.Cp; ++i) {
                const int c = (int)(i % A.Cp);
                if (A.also_zero) A.also_zero[i] = 0.0f;
                A.grid_cl[i] = c < A.C ? A.coeff[0][(long long)c * nvox + i / A.Cp] : 0.0f;
            }
        } else {
            for (int l = 1; l < A.n_coeff; ++l) {
                const bool last = l == A.n_coeff - 1;
                const long long total = (long long)A.t[l][0] * A.t[l][1] * A.t[l][2] * (last ? A.Cp : A.C);
                for (long long i = 0; i < total; ++i) synth_elem<NT>(A, l, i);
            }
        }
    }
}

extern "C" int lfgc_step_glue_host(const lfgc_wavelet_desc* w, int Cp, float* const* coeff, float* const* grad_coeff,
                                   float* scratch, const float* grad_grid_cl, float* grid_cl, float* also_zero, float* p,
                                   const float* g, float* m, float* v, int64_t n, const float* lr, int32_t* step_count,
                                   double beta1, double beta2, double eps, double grad_scale, int phases) {
    if ((phases & 7) == 0) return LFGC_OK;
    glue::Args A;
    const int rc = glue_fill_args(A, w, Cp, coeff, grad_coeff, scratch, grad_grid_cl, grid_cl, also_zero, p, g, m, v, n, lr,
                                  step_count, beta1, beta2, eps, grad_scale, phases);
    if (rc) return rc;
    if (w->n_taps == 2 || w->n_coeff == 1) glue_run_host<2>(A);
    else glue_run_host<4>(A);
    return LFGC_OK;
}
#endif
