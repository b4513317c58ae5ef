// Flat-buffer optimiser pieces: Adam (torch.optim.Adam semantics, training/training.py:199,232) and the
// sample-independent regulariser gradients of SmallifyLoss (model/Smallify_Dropout.py:22-40).
#include "lfgc_common.cuh"

namespace lfgc {

// step_ptr[0] = optimiser steps taken so far, step_ptr[1] = ticket counter (must start at 0).  Every block reads the
// step before it takes its ticket and the block drawing the last ticket publishes step + 1, so the whole update is
// one launch and stays correct under CUDA-graph replay.
// Optional SmallifyLoss terms (model/Smallify_Dropout.py:22-40) folded into the update: + 2 w2 p on [l2_begin, l2_end)
// (the wavelet coefficients), + w1 sign(p) on [l1_begin, l1_end) (the mask parameters).  Empty ranges: plain Adam.
struct AdamReg {
    long long l2_begin, l2_end, l1_begin, l1_end;
    float w2x2, w1;
    // Smallify bookkeeping of the mask range, riding along (all optional): the sign-variance tracker update
    // (Smallify_Dropout.py:103-112) with the parameter value this step's forward used (i.e. before the update below),
    // and clearing the gradient element, which the synthesis adjoint accumulates into with atomics
    float* ema;
    float* emavar;
    float momentum;
    int zero_l1_grad;
};
__device__ __forceinline__ void adam_reg_side(float* g, float p_old, long long i, const AdamReg& r) {
    if (i < r.l1_begin || i >= r.l1_end) return;
    if (r.ema) {
        const long long j = i - r.l1_begin;
        const float sgn = (p_old > 0.0f) ? 1.0f : ((p_old < 0.0f) ? -1.0f : 0.0f);
        const float e = r.ema[j];
        const float phi = __fsub_rn(sgn, e);
        r.ema[j] = __fadd_rn(e, __fmul_rn(r.momentum, phi));
        r.emavar[j] = __fmul_rn(__fsub_rn(1.0f, r.momentum), __fadd_rn(r.emavar[j], __fmul_rn(r.momentum, __fmul_rn(phi, phi))));
    }
    if (r.zero_l1_grad) g[i] = 0.0f;
}
__device__ __forceinline__ float adam_reg_grad(float g, float p, long long i, const AdamReg& r) {
    if (i >= r.l2_begin && i < r.l2_end) g = fmaf(r.w2x2, p, g);
    if (i >= r.l1_begin && i < r.l1_end) g += r.w1 * ((p > 0.0f) ? 1.0f : ((p < 0.0f) ? -1.0f : 0.0f));
    return g;
}

__global__ void __launch_bounds__(256, 4) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, const float* __restrict__ lr_ptr,
                            int32_t* __restrict__ step_ptr, const AdamCoef c, const AdamReg reg) {
    LFGC_PDL_PROLOGUE();
    __shared__ float s_step_size, s_bc2_sqrt;
    __shared__ int s_step;
    // this thread's first batch goes out before anything else: the loads overlap thread 0's scalar work below
    float4 pre_p = make_float4(0.f, 0.f, 0.f, 0.f), pre_m = pre_p, pre_v = pre_p, pre_g = pre_p;
    {
        const bool vec0 = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                            reinterpret_cast<uintptr_t>(v)) & 15) == 0;
        const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (vec0 && i0 < (n >> 2)) {
            pre_p = reinterpret_cast<const float4*>(p)[i0];
            pre_m = reinterpret_cast<const float4*>(m)[i0];
            pre_v = reinterpret_cast<const float4*>(v)[i0];
            pre_g = reinterpret_cast<const float4*>(g)[i0];
        }
    }
    if (threadIdx.x == 0) {
        const int step = *reinterpret_cast<volatile int32_t*>(step_ptr) + 1;
        s_step = step;
        // bias corrections in fp32 from fl32(log(beta)) (see AdamCoef): ~1e-7 relative, without the cancellation of
        // 1 - pow() and without the slow fp64 pipe (one thread per block sits on this latency)
        float step_size, bc2_sqrt;
        adam_step_scalars(c, step, *lr_ptr, step_size, bc2_sqrt);
        s_step_size = step_size;
        s_bc2_sqrt = bc2_sqrt;
    }
    __syncthreads();
    // Grid-stride over 128-bit batches (four parameters) when the buffers allow it -- the flat buffers of FastTrainer do --
    // with the NEXT batch's four loads issued before the current one is processed: 28 B of HBM traffic per parameter is all
    // this kernel is, so what matters is bytes in flight per SM (one batch per short-lived CTA, the first version, reached
    // 2.7 TB/s on the 8.4 M parameters of the C32/G64 grid).
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    const int64_t nvec = vec ? (n >> 2) : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
    const bool side = reg.l1_end > reg.l1_begin;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float4 p4 = pre_p, m4 = pre_m, v4 = pre_v, g4 = pre_g;
    while (i < nvec) {
        const int64_t inext = i + stride;
        float4 pn, mn, vn, gn;
        if (inext < nvec) {
            pn = reinterpret_cast<const float4*>(p)[inext];
            mn = reinterpret_cast<const float4*>(m)[inext];
            vn = reinterpret_cast<const float4*>(v)[inext];
            gn = reinterpret_cast<const float4*>(g)[inext];
        }
        const int64_t i4 = i << 2;
        if (side) {
            adam_reg_side(g, p4.x, i4, reg);
            adam_reg_side(g, p4.y, i4 + 1, reg);
            adam_reg_side(g, p4.z, i4 + 2, reg);
            adam_reg_side(g, p4.w, i4 + 3, reg);
        }
        adam_update(p4.x, adam_reg_grad(g4.x, p4.x, i4, reg), m4.x, v4.x, c, step_size, bc2_sqrt);
        adam_update(p4.y, adam_reg_grad(g4.y, p4.y, i4 + 1, reg), m4.y, v4.y, c, step_size, bc2_sqrt);
        adam_update(p4.z, adam_reg_grad(g4.z, p4.z, i4 + 2, reg), m4.z, v4.z, c, step_size, bc2_sqrt);
        adam_update(p4.w, adam_reg_grad(g4.w, p4.w, i4 + 3, reg), m4.w, v4.w, c, step_size, bc2_sqrt);
        reinterpret_cast<float4*>(p)[i] = p4;
        reinterpret_cast<float4*>(m)[i] = m4;
        reinterpret_cast<float4*>(v)[i] = v4;
        i = inext;
        p4 = pn; m4 = mn; v4 = vn; g4 = gn;
    }
    // the elements the 128-bit batches do not cover (n % 4, or everything when a buffer is not 16-byte aligned)
    for (int64_t e = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        float pi = p[e], mi = m[e], vi = v[e];
        const float gi = g[e];
        adam_reg_side(g, pi, e, reg);
        adam_update(pi, adam_reg_grad(gi, pi, e, reg), mi, vi, c, step_size, bc2_sqrt);
        p[e] = pi;
        m[e] = mi;
        v[e] = vi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int ticket = atomicAdd(step_ptr + 1, 1);
        if (ticket == (int)gridDim.x - 1) {
            step_ptr[1] = 0;
            __threadfence();
            step_ptr[0] = s_step;
        }
    }
}

// Data-parallel gradient sum WITHOUT a collective: out[i] = sum over the ranks (in rank order, so every rank forms
// bit-identical sums) of their gradient buffers, read straight over NVLink from peer memory (torch symmetric memory
// provides the mapping), behind a barrier that lives INSIDE the kernel: every rank's kernel announces its epoch to all
// ranks' flag arrays with a system-scope release store and waits until every rank has announced the same epoch.  A rank
// reaches this point only after its own per-sample kernel (same stream) has completed, so once the wait is over every
// rank's buffer of this step is complete and visible.  Epochs only grow: nothing to reset; a lost peer traps after ~2 s
// instead of hanging the device.  The caller double-buffers the sources by step parity (a rank may only overwrite a buffer
// its peers read one epoch later), which is why ONE barrier per step is enough; `zero` (the other parity's local buffer)
// is cleared here for the next step.  For the 0.23 MB message of the shipped configurations this replaces an NCCL
// all-reduce whose cost is all latency (+18 us per step at 2 GPUs, +36 us at 8, SCALE_r01).
struct PeerSumArgs {
    const float* src[LFGC_MAX_PEERS];
    int* flags[LFGC_MAX_PEERS];
    int n_srcs, rank;
    int* epoch;
    float* out;
    float* zero;
    long long n;
    int announced;   // this epoch's flags were already stored by the per-sample kernel (lfgc_peer_announce)
};

#ifdef LFGC_PHASE_TIMING
__device__ unsigned long long g_peer_ns[8];   // [0] barrier wait, [1] remote reads + sum, [2] launches, [3] entry skew vs the train kernel
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif
__global__ void __launch_bounds__(512) peer_sum_kernel(const __grid_constant__ PeerSumArgs A) {
    LFGC_PDL_PROLOGUE();
    const int tid = threadIdx.x;
#ifdef LFGC_PHASE_TIMING
    const unsigned long long t_in = gtimer();
#endif
    const int e = *reinterpret_cast<volatile int*>(A.epoch) + 1;
    if (tid < A.n_srcs) {
        if (blockIdx.x == 0 && !A.announced) {
            // everything this rank contributes was written by EARLIER kernels of this stream (complete and visible at
            // device scope, i.e. in this GPU's L2, which is where peers read it): the release store alone orders it
            asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(A.flags[tid] + A.rank), "r"(e) : "memory");
        }
        const int* mine = A.flags[A.rank] + tid;
        const long long t0 = clock64();
        for (;;) {
            int seen;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
            if (seen >= e) break;
            if (clock64() - t0 > 4000000000ll) __trap();
        }
    }
    __syncthreads();
#ifdef LFGC_PHASE_TIMING
    const unsigned long long t_bar = gtimer();
#endif
    const long long n4 = A.n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < n4; i += (long long)gridDim.x * blockDim.x) {
        // peer data changes every step: never from L1.  Four sources' loads are in flight together (NVLink round trips
        // overlap), the sum stays in rank order
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r0 = 0; r0 < A.n_srcs; r0 += 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r0 + u < A.n_srcs) v[u] = __ldcv(reinterpret_cast<const float4*>(A.src[r0 + u]) + i);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r0 + u < A.n_srcs) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        reinterpret_cast<float4*>(A.out)[i] = acc;
        if (A.zero) reinterpret_cast<float4*>(A.zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // publish the new epoch once every CTA has read the old one (ticket in epoch[1])
    __syncthreads();
#ifdef LFGC_PHASE_TIMING
    if (tid == 0 && blockIdx.x == 0) {
        const unsigned long long t_end = gtimer();
        atomicAdd(&g_peer_ns[0], t_bar - t_in);
        atomicAdd(&g_peer_ns[1], t_end - t_bar);
        atomicAdd(&g_peer_ns[2], 1ull);
        g_peer_ns[4] = t_in;     // last launch: entry and exit of CTA 0
        g_peer_ns[5] = t_end;
    }
#endif
    if (tid == 0) {
        __threadfence();
        const int ticket = atomicAdd(A.epoch + 1, 1);
        if (ticket == (int)gridDim.x - 1) {
            A.epoch[1] = 0;
            __threadfence();
            A.epoch[0] = e;
        }
    }
}

__global__ void add_l2_grad_kernel(float* __restrict__ g, const float* __restrict__ p, int64_t n, float w2) {
    LFGC_PDL_PROLOGUE();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g[i] = fmaf(w2, p[i], g[i]);
}
__global__ void add_l1_grad_kernel(float* __restrict__ g, const float* __restrict__ p, int64_t n, float w) {
    LFGC_PDL_PROLOGUE();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float x = p[i];
        g[i] += w * ((x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f));
    }
}

// Gradient of the KL term of VariationalDropoutLoss (model/Variational_Dropout_Layer.py:54-69,115-122) with respect to
// the mask parameters, plus the per-step ramp of its weight.  The mask section of the flat parameter buffer holds, per
// mask layer i, [log_thetas_i (n_i) | log_var_i (n_i)]; seg.end[i] is the end (in mask elements, cumulative n_i) of
// layer i.  DKL = sum(-k1 sigmoid(k2 + k3 la) + 0.5 softplus(-la) + k1), la = log_var - 2 log_thetas.
// w_dkl[2] is a ping-pong pair indexed by the parity of the optimiser step counter: every thread reads the current
// weight, applies the ramp locally and one thread publishes the ramped value for the next step, so the whole thing is
// one launch and graph-replayable.
struct DklSegments {
    int n_layers;
    long long end[LFGC_MAX_LEVELS];
};

__global__ void variational_dkl_grad_kernel(const float* __restrict__ mask_p, float* __restrict__ mask_g,
                                            const __grid_constant__ DklSegments seg, double* __restrict__ w_dkl,
                                            const int32_t* __restrict__ step_ptr, double ramp, double w_max,
                                            float scale) {
    LFGC_PDL_PROLOGUE();
    const int step = step_ptr ? *step_ptr : 0;
    const double w_old = w_dkl[step & 1];
    const double w_new = w_old < w_max ? w_old * ramp : w_old;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) w_dkl[(step + 1) & 1] = w_new;
    const long long total = seg.end[seg.n_layers - 1];
    if (i >= total) return;
    int layer = 0;
    while (i >= seg.end[layer]) ++layer;
    const long long begin = layer ? seg.end[layer - 1] : 0;
    const long long n = seg.end[layer] - begin;
    const long long lt = 2 * begin + (i - begin), lv = lt + n;
    const float la = mask_p[lv] - 2.0f * mask_p[lt];
    const float k1 = 0.63576f, k2 = 1.87320f, k3 = 1.48695f;
    const float su = 1.0f / (1.0f + expf(-(k2 + k3 * la)));
    const float sm = 1.0f / (1.0f + expf(la));                 // sigmoid(-la)
    const float d = -k1 * k3 * su * (1.0f - su) - 0.5f * sm;   // d DKL / d la
    const float g = (float)w_new * scale * d;
    mask_g[lv] += g;
    mask_g[lt] -= 2.0f * g;
}

// All live variational mask layers in ONE launch each way (the per-layer lfgc_mask_multiplier / lfgc_mask_param_grad calls
// plus their copies were ~20 graph nodes of the variational step).  Position q of the concatenated masks belongs to layer
// `layer`; its parameters sit in the flat mask section at lt = 2 begin + (q - begin) (log_thetas) and lv = lt + n (log_var).
__device__ __forceinline__ void dkl_locate(const DklSegments& seg, long long q, long long& lt, long long& lv) {
    int layer = 0;
    while (q >= seg.end[layer]) ++layer;
    const long long begin = layer ? seg.end[layer - 1] : 0;
    const long long n = seg.end[layer] - begin;
    lt = 2 * begin + (q - begin);
    lv = lt + n;
}
// mult = exp(log_thetas) + exp(log_var / 2) xi   (Variational_Dropout_Layer.py:104-108); zero_out (nullable) is cleared
__global__ void variational_multiplier_kernel(const float* __restrict__ mask_p, const float* __restrict__ noise,
                                              const __grid_constant__ DklSegments seg, float* __restrict__ mult,
                                              float* __restrict__ zero_out) {
    LFGC_PDL_PROLOGUE();
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= seg.end[seg.n_layers - 1]) return;
    long long lt, lv;
    dkl_locate(seg, q, lt, lv);
    mult[q] = __fadd_rn(expf(mask_p[lt]), __fmul_rn(expf(mask_p[lv] / 2.0f), noise[q]));
    if (zero_out) zero_out[q] = 0.0f;
}
// d loss / d(log_thetas, log_var) from d loss / d mult (autograd of the line above), written into the flat gradient
__global__ void variational_param_grad_kernel(const float* __restrict__ mask_p, const float* __restrict__ noise,
                                              const float* __restrict__ gmult, const __grid_constant__ DklSegments seg,
                                              float* __restrict__ mask_g) {
    LFGC_PDL_PROLOGUE();
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= seg.end[seg.n_layers - 1]) return;
    long long lt, lv;
    dkl_locate(seg, q, lt, lv);
    const float gm = gmult[q];
    mask_g[lt] = gm * expf(mask_p[lt]);
    mask_g[lv] = gm * (0.5f * expf(mask_p[lv] / 2.0f) * noise[q]);
}

}  // namespace lfgc

using namespace lfgc;

extern "C" int lfgc_variational_dkl_grad(const float* mask_params, float* mask_grads, int n_layers,
                                         const int64_t* layer_sizes, double* w_dkl, const int32_t* step_count,
                                         double ramp, double w_max, float scale, void* stream) {
    if (!mask_params || !mask_grads || !layer_sizes || !w_dkl || n_layers < 1 || n_layers > LFGC_MAX_LEVELS)
        return fail(LFGC_E_INVALID, "variational_dkl_grad: bad arguments");
    DklSegments seg;
    seg.n_layers = n_layers;
    long long acc = 0;
    for (int i = 0; i < n_layers; ++i) {
        if (layer_sizes[i] < 0) return fail(LFGC_E_INVALID, "variational_dkl_grad: negative layer size");
        acc += layer_sizes[i];
        seg.end[i] = acc;
    }
    const long long blocks = acc == 0 ? 1 : (acc + 255) / 256;
    (void)launch_pdl(variational_dkl_grad_kernel, dim3((unsigned)blocks), dim3(256), (size_t)(0), (cudaStream_t)stream, mask_params, mask_grads, seg, w_dkl,
                                                                                  step_count, ramp, w_max, scale);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}


static int launch_adam(float* p, float* g, float* m, float* v, int64_t n, const float* lr, int32_t* step_count,
                       double beta1, double beta2, double eps, double grad_scale, const AdamReg& reg, void* stream) {
    if (!p || !g || !m || !v || !lr || !step_count || n < 0) return fail(LFGC_E_INVALID, "adam: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = n == 0 ? 1 : (n + 1023) / 1024;   // 256 threads x 4 parameters per batch ...
    const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 4;
    if (blocks > cap) blocks = cap;                          // ... grid-stride beyond one wave of 4 resident CTAs per SM
    (void)launch_pdl(adam_kernel, dim3((unsigned)blocks), dim3(256), (size_t)(0), st, p, g, m, v, n, lr, step_count,
                     make_adam_coef(beta1, beta2, eps, grad_scale), reg);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

static int fill_segments(DklSegments& seg, int n_layers, const int64_t* layer_sizes, long long& total) {
    if (!layer_sizes || n_layers < 1 || n_layers > LFGC_MAX_LEVELS) return fail(LFGC_E_INVALID, "variational masks: bad layer list");
    seg.n_layers = n_layers;
    long long acc = 0;
    for (int i = 0; i < n_layers; ++i) {
        if (layer_sizes[i] < 0) return fail(LFGC_E_INVALID, "variational masks: negative layer size");
        acc += layer_sizes[i];
        seg.end[i] = acc;
    }
    total = acc;
    return LFGC_OK;
}

extern "C" int lfgc_variational_multiplier(const float* mask_params, const float* noise, int n_layers,
                                           const int64_t* layer_sizes, float* mult_out, float* zero_out, void* stream) {
    if (!mask_params || !noise || !mult_out) return fail(LFGC_E_INVALID, "variational_multiplier: null pointer");
    DklSegments seg;
    long long total = 0;
    const int rc = fill_segments(seg, n_layers, layer_sizes, total);
    if (rc) return rc;
    if (total == 0) return LFGC_OK;
    (void)launch_pdl(variational_multiplier_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)0, (cudaStream_t)stream,
                     mask_params, noise, seg, mult_out, zero_out);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

extern "C" int lfgc_variational_param_grad(const float* mask_params, const float* noise, const float* gmult, int n_layers,
                                           const int64_t* layer_sizes, float* mask_grads, void* stream) {
    if (!mask_params || !noise || !gmult || !mask_grads) return fail(LFGC_E_INVALID, "variational_param_grad: null pointer");
    DklSegments seg;
    long long total = 0;
    const int rc = fill_segments(seg, n_layers, layer_sizes, total);
    if (rc) return rc;
    if (total == 0) return LFGC_OK;
    (void)launch_pdl(variational_param_grad_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)0, (cudaStream_t)stream,
                     mask_params, noise, gmult, seg, mask_grads);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

extern "C" int lfgc_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* lr, int32_t* step_count,
                         double beta1, double beta2, double eps, double grad_scale, void* stream) {
    const AdamReg none = {0, 0, 0, 0, 0.0f, 0.0f, nullptr, nullptr, 0.0f, 0};
    return launch_adam(p, const_cast<float*>(g), m, v, n, lr, step_count, beta1, beta2, eps, grad_scale, none, stream);
}

extern "C" int lfgc_adam_reg(float* p, float* g, float* m, float* v, int64_t n, const float* lr, int32_t* step_count,
                             double beta1, double beta2, double eps, double grad_scale, int64_t l2_begin, int64_t l2_end,
                             double weight_l2, int64_t l1_begin, int64_t l1_end, double weight_l1, float* ema,
                             float* emavar, float momentum, int zero_l1_grad, void* stream) {
    if ((ema == nullptr) != (emavar == nullptr)) return fail(LFGC_E_INVALID, "adam_reg: ema and emavar go together");
    if (l2_begin < 0 || l2_end > n || l1_begin < 0 || l1_end > n) return fail(LFGC_E_INVALID, "adam_reg: range outside the buffer");
    AdamReg reg;
    reg.l2_begin = l2_begin;
    reg.l2_end = l2_end > l2_begin ? l2_end : l2_begin;
    reg.l1_begin = l1_begin;
    reg.l1_end = l1_end > l1_begin ? l1_end : l1_begin;
    reg.w2x2 = (float)(2.0 * weight_l2);
    reg.w1 = (float)weight_l1;
    reg.ema = ema;
    reg.emavar = emavar;
    reg.momentum = momentum;
    reg.zero_l1_grad = zero_l1_grad;
    return launch_adam(p, g, m, v, n, lr, step_count, beta1, beta2, eps, grad_scale, reg, stream);
}

#ifdef LFGC_PHASE_TIMING
extern "C" int lfgc_peer_timing(unsigned long long* out8, int reset) {
    cudaDeviceSynchronize();
    if (out8) cudaMemcpyFromSymbol(out8, g_peer_ns, sizeof(unsigned long long) * 8);
    if (reset) {
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(g_peer_ns, z, sizeof(z));
    }
    return 0;
}
#endif

extern "C" int lfgc_peer_sum(const float* const* srcs, int32_t* const* flags, int n_srcs, int rank, int32_t* epoch, float* out,
                             float* zero, int64_t n, int announced, void* stream) {
    if (!srcs || !flags || !epoch || !out || n < 0 || (n & 3)) return fail(LFGC_E_INVALID, "peer_sum: bad arguments (n must be a multiple of 4)");
    if (n_srcs < 1 || n_srcs > LFGC_MAX_PEERS || rank < 0 || rank >= n_srcs) return fail(LFGC_E_UNSUPPORTED, "peer_sum: %d sources, rank %d", n_srcs, rank);
    PeerSumArgs A = {};
    for (int r = 0; r < n_srcs; ++r) {
        if (!srcs[r] || !flags[r]) return fail(LFGC_E_INVALID, "peer_sum: pointer of rank %d is null", r);
        A.src[r] = srcs[r];
        A.flags[r] = flags[r];
    }
    A.n_srcs = n_srcs;
    A.rank = rank;
    A.epoch = epoch;
    A.out = out;
    A.zero = zero;
    A.n = n;
    A.announced = announced ? 1 : 0;
    int blocks = (int)((n / 4 + 511) / 512);
    if (blocks < 1) blocks = 1;
    if (blocks > 2 * sm_count()) blocks = 2 * sm_count();   // all CTAs must be co-resident while they wait on the flags
    (void)launch_pdl(peer_sum_kernel, dim3((unsigned)blocks), dim3(512), (size_t)0, (cudaStream_t)stream, A);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

extern "C" int lfgc_add_l2_grad(float* g, const float* p, int64_t n, float weight, void* stream) {
    if (!g || !p || n < 0) return fail(LFGC_E_INVALID, "add_l2_grad: bad arguments");
    if (n == 0) return LFGC_OK;
    (void)launch_pdl(add_l2_grad_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, g, p, n, 2.0f * weight);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

extern "C" int lfgc_add_l1_grad(float* g, const float* p, int64_t n, float weight, void* stream) {
    if (!g || !p || n < 0) return fail(LFGC_E_INVALID, "add_l1_grad: bad arguments");
    if (n == 0) return LFGC_OK;
    (void)launch_pdl(add_l1_grad_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, g, p, n, weight);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}
