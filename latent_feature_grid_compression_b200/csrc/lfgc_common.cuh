// Shared host/device helpers for liblfgc (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "lfgc.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "liblfgc is written for sm_100a (B200) only"
#endif

namespace lfgc {

// ---- error plumbing ---------------------------------------------------------------------------------------------
char* last_error_buffer();  // thread-local, defined in api.cu
void count_launch();        // host-side counter of kernel launches issued by this library (api.cu)

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define LFGC_CUDA_OK(expr)                                                                          \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return ::lfgc::fail(LFGC_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                __FILE__, __LINE__);                                                \
    } while (0)

#define LFGC_LAUNCH_OK()                                                                           \
    do {                                                                                            \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e != cudaSuccess)                                                                      \
            return ::lfgc::fail(LFGC_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                                __FILE__, __LINE__);                                                \
        ::lfgc::count_launch();                                                                     \
    } while (0)

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// The kernels of one training step form a chain of short dependent launches (synthesis levels -> fused training kernel
// -> partial reduction -> adjoint levels -> Adam).  Every kernel on that chain starts with LFGC_PDL_PROLOGUE():
// griddepcontrol.wait (all prerequisite grids have completed and their writes are visible) followed by
// griddepcontrol.launch_dependents, and is launched with the programmatic-stream-serialization attribute, so the
// launch latency and block scheduling of kernel N+1 overlap kernel N instead of following it.  The trigger comes
// after the wait on purpose: completion is then transitive along the chain.  Launched without the attribute (the
// default: inside a CUDA graph the edges gained nothing, see api.cu) the two instructions are no-ops.
bool pdl_enabled();        // api.cu: LFGC_PDL=1 environment switch (off by default, see there)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// torch.optim.Adam coefficients as torch forms them: the betas are Python doubles there, so 1 - beta and the bias
// corrections come from the fp64 values and are only then rounded to fp32 (1 - fl32(0.999) would be 4.7e-5 off).
struct AdamCoef {
    float b1, b2;          // fl32(beta)
    float omb1, omb2;      // fl32(1 - beta)
    float logb1, logb2;    // fl32(log(beta)): bias correction 1 - beta^step = -expm1(step * log(beta))
    float eps, gscale;
};
AdamCoef make_adam_coef(double beta1, double beta2, double eps, double grad_scale);   // api.cu

int sm_count();            // cached per device
int max_smem_optin();      // cached per device (bytes)

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---- device helpers ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

#define LFGC_PDL_PROLOGUE()                                      \
    do {                                                         \
        asm volatile("griddepcontrol.wait;" ::: "memory");       \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
    } while (0)

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    // vectorised fp32 reduction straight into L2 (sm_90+): one 16-byte atomic per 4 channels
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// One element of torch.optim.Adam (training/training.py:199,232; torch/optim/adam.py _single_tensor_adam):
//   exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2);
//   denom = exp_avg_sq.sqrt() / sqrt(bias_correction2) + eps; param.addcdiv_(exp_avg, denom, value=-lr / bias_correction1)
__host__ __device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamCoef& c,
                                                     float step_size, float bc2_sqrt) {
    const float gi = g * c.gscale;
    float mi = m, vi = v;
    mi = mi + (gi - mi) * c.omb1;
    vi = vi * c.b2 + c.omb2 * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + c.eps;
    p = p - step_size * (mi / denom);
    m = mi;
    v = vi;
}
// step_size = lr / (1 - beta1^step), bc2_sqrt = sqrt(1 - beta2^step)
__host__ __device__ __forceinline__ void adam_step_scalars(const AdamCoef& c, int step, float lr, float& step_size,
                                                           float& bc2_sqrt) {
    const float bc1 = -expm1f((float)step * c.logb1);
    const float bc2 = -expm1f((float)step * c.logb2);
    step_size = lr / bc1;
    bc2_sqrt = sqrtf(bc2);
}

// Branch-free fp32 sin/cos: Cody-Waite reduction by pi/2 (three-term split, exact products through FMA) followed
// by the classic minimax polynomials on [-pi/4, pi/4].  Max error ~1 ulp for |x| < 1e4 (beyond that the 3-term
// reduction slowly loses bits; MLP pre-activations and embedding arguments are orders of magnitude smaller).
// About 20 instructions for both values and no divergent slow path, unlike sincosf.
__device__ __forceinline__ void sincos_cw(float x, float& s, float& c) {
    // n = rint(x * 2/pi) by the magic-number trick (two FADDs on the FMA pipe instead of FRND + F2I on the
    // quarter-rate conversion unit); the low mantissa bits of the biased sum are the quadrant.  Valid for
    // |x * 2/pi| < 2^22, far beyond the 1e4 the reduction below is accurate for.
    const float biased = __fadd_rn(__fmul_rn(x, 0.636619772f), 12582912.0f);  // 1.5 * 2^23
    const int q = __float_as_int(biased);
    const float n = __fsub_rn(biased, 12582912.0f);
    float r = fmaf(n, -1.57079601e+00f, x);
    r = fmaf(n, -3.13916473e-07f, r);
    r = fmaf(n, -5.39030253e-15f, r);
    const float r2 = r * r;
    float ps = fmaf(r2, -1.95152959e-4f, 8.33216087e-3f);
    ps = fmaf(ps, r2, -1.66666546e-1f);
    ps = fmaf(ps * r2, r, r);                 // sin(r)
    float pc = fmaf(r2, 2.44331571e-5f, -1.38873163e-3f);
    pc = fmaf(pc, r2, 4.16666457e-2f);
    pc = fmaf(pc, r2, -0.5f);
    pc = fmaf(pc, r2, 1.0f);                  // cos(r)
    const float sv = (q & 1) ? pc : ps;
    const float cv = (q & 1) ? ps : pc;
    s = (q & 2) ? -sv : sv;
    c = ((q + 1) & 2) ? -cv : cv;
}

// SnakeAlt(x) = 0.5 x + sin^2 x and its derivative 0.5 + sin 2x = 0.5 + 2 sin x cos x
// (model/Feature_Grid_Model.py:12-13).
#ifndef LFGC_SNAKE_MUFU
#define LFGC_SNAKE_MUFU 0
#endif
#if LFGC_SNAKE_MUFU
// sin^2 z = (1 - cos 2z) / 2 and sin 2z on the special-function unit after an exact reduction of z modulo pi
// (2r in [-pi, pi] is the interval MUFU.SIN / MUFU.COS are specified on: max abs error 2^-21.4 / 2^-21.2).
// 11 instructions, 2 of them on the otherwise idle XU pipe, instead of ~24 FMA-pipe instructions.
__device__ __forceinline__ void snake_cos_sin_2z(float z, float& c2, float& s2) {
    const float biased = __fadd_rn(__fmul_rn(z, 0.318309886f), 12582912.0f);   // rint(z / pi)
    const float n = __fsub_rn(biased, 12582912.0f);
    float r = fmaf(n, -3.14159274e+00f, z);
    r = fmaf(n, 8.74227766e-08f, r);
    const float w = r + r;
    c2 = __cosf(w);
    s2 = __sinf(w);
}
__device__ __forceinline__ float snake_precise(float z) {
    float c2, s2;
    snake_cos_sin_2z(z, c2, s2);
    return fmaf(-0.5f, c2, fmaf(0.5f, z, 0.5f));
}
__device__ __forceinline__ void snake_and_grad_precise(float z, float& h, float& g) {
    float c2, s2;
    snake_cos_sin_2z(z, c2, s2);
    h = fmaf(-0.5f, c2, fmaf(0.5f, z, 0.5f));
    g = s2 + 0.5f;
}
#else
__device__ __forceinline__ float snake_precise(float z) {
    float s, c;
    sincos_cw(z, s, c);
    return fmaf(s, s, 0.5f * z);
}
__device__ __forceinline__ void snake_and_grad_precise(float z, float& h, float& g) {
    float s, c;
    sincos_cw(z, s, c);
    h = fmaf(s, s, 0.5f * z);
    g = fmaf(2.0f * s, c, 0.5f);
}
#endif

// SnakeAlt VALUE only (forward / reconstruction): 0.5 z + sin^2 z = 0.5 z + 0.5 - 0.5 cos 2z with ONE polynomial.
// 2z is reduced modulo pi (two-term Cody-Waite, exact products through FMA) to r in [-pi/2, pi/2], where cos r is a
// degree-5 polynomial in r^2 (least-squares fit on Chebyshev nodes, max abs error 6.2e-8 in fp32 Horner form); the sign
// (-1)^n goes into the coefficient -0.5 by flipping its sign bit.  15 instructions instead of ~26 for the sin AND cos
// polynomials plus quadrant selects of sincos_cw, and no select at all.  Valid for |z| < 1e4 like sincos_cw.
#define LFGC_SNAKE_C2 4.16666343808e-2f
#define LFGC_SNAKE_C3 -1.38883292675e-3f
#define LFGC_SNAKE_C4 2.47584812314e-5f
#define LFGC_SNAKE_C5 -2.60215898606e-7f
__device__ __forceinline__ float snake_value_fast(float z) {
    const float w = z + z;
    const float biased = fmaf(w, 0.318309886f, 12582912.0f);   // rint(2z / pi) by the magic-number trick
    const float n = biased - 12582912.0f;
    float r = fmaf(n, -3.14159274f, w);
    r = fmaf(n, 8.74227766e-08f, r);
    const float u = r * r;
    float p = fmaf(u, LFGC_SNAKE_C5, LFGC_SNAKE_C4);
    p = fmaf(p, u, LFGC_SNAKE_C3);
    p = fmaf(p, u, LFGC_SNAKE_C2);
    p = fmaf(p, u, -0.5f);
    p = fmaf(p, u, 1.0f);                                       // cos r
    const float sgn = __int_as_float(0xBF000000 ^ (__float_as_int(biased) << 31));   // -0.5 (-1)^n
    return fmaf(sgn, p, fmaf(0.5f, z, 0.5f));
}

// The same for two values at once with packed fp32 arithmetic (fma.rn.f32x2, sm_100+: one issue slot per pair).
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ void snake_value_fast2(float z0, float z1, float& h0, float& h1) {
    const unsigned long long z = pack2(z0, z1);
    const unsigned long long one = pack2(1.0f, 1.0f), half = pack2(0.5f, 0.5f), magic = pack2(12582912.0f, 12582912.0f);
    const unsigned long long w = fma2(z, one, z);                                        // 2 z
    const unsigned long long biased = fma2(w, pack2(0.318309886f, 0.318309886f), magic);
    const unsigned long long n = fma2(biased, one, pack2(-12582912.0f, -12582912.0f));
    unsigned long long r = fma2(n, pack2(-3.14159274f, -3.14159274f), w);
    r = fma2(n, pack2(8.74227766e-08f, 8.74227766e-08f), r);
    const unsigned long long u = fma2(r, r, pack2(0.0f, 0.0f));
    unsigned long long p = fma2(u, pack2(LFGC_SNAKE_C5, LFGC_SNAKE_C5), pack2(LFGC_SNAKE_C4, LFGC_SNAKE_C4));
    p = fma2(p, u, pack2(LFGC_SNAKE_C3, LFGC_SNAKE_C3));
    p = fma2(p, u, pack2(LFGC_SNAKE_C2, LFGC_SNAKE_C2));
    p = fma2(p, u, pack2(-0.5f, -0.5f));
    p = fma2(p, u, one);
    float b0, b1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(b0), "=f"(b1) : "l"(biased));
    const unsigned long long sgn = pack2(__int_as_float(0xBF000000 ^ (__float_as_int(b0) << 31)),
                                         __int_as_float(0xBF000000 ^ (__float_as_int(b1) << 31)));
    const unsigned long long h = fma2(sgn, p, fma2(half, z, half));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(h0), "=f"(h1) : "l"(h));
}

// SnakeAlt AND its derivative for two values at once (training kernel epilogue): h = 0.5 z + 0.5 - 0.5 cos 2z,
// g = 0.5 + sin 2z, with 2z reduced modulo pi as above and sin r on [-pi/2, pi/2] as r + r u (S1 + u (S2 + u (S3 + u S4)))
// (weighted least-squares fit, max abs error 1.1e-7 in fp32 Horner form).  19 packed + 6 scalar instructions per PAIR
// against ~29 scalar instructions per VALUE for sincos_cw + quadrant selects.
#define LFGC_SNAKE_S1 -0.16666653752326965f
#define LFGC_SNAKE_S2 0.008332948200404644f
#define LFGC_SNAKE_S3 -0.00019802302995231003f
#define LFGC_SNAKE_S4 2.5916351660271175e-06f
__device__ __forceinline__ void snake_and_grad_fast2(float z0, float z1, float& h0, float& h1, float& g0, float& g1) {
    const unsigned long long z = pack2(z0, z1);
    const unsigned long long one = pack2(1.0f, 1.0f), half = pack2(0.5f, 0.5f), magic = pack2(12582912.0f, 12582912.0f);
    const unsigned long long zero = pack2(0.0f, 0.0f);
    const unsigned long long w = fma2(z, one, z);                                        // 2 z
    const unsigned long long biased = fma2(w, pack2(0.318309886f, 0.318309886f), magic);
    const unsigned long long n = fma2(biased, one, pack2(-12582912.0f, -12582912.0f));
    unsigned long long r = fma2(n, pack2(-3.14159274f, -3.14159274f), w);
    r = fma2(n, pack2(8.74227766e-08f, 8.74227766e-08f), r);
    const unsigned long long u = fma2(r, r, zero);
    unsigned long long p = fma2(u, pack2(LFGC_SNAKE_C5, LFGC_SNAKE_C5), pack2(LFGC_SNAKE_C4, LFGC_SNAKE_C4));
    p = fma2(p, u, pack2(LFGC_SNAKE_C3, LFGC_SNAKE_C3));
    p = fma2(p, u, pack2(LFGC_SNAKE_C2, LFGC_SNAKE_C2));
    p = fma2(p, u, pack2(-0.5f, -0.5f));
    p = fma2(p, u, one);                                                                 // cos r
    unsigned long long t = fma2(u, pack2(LFGC_SNAKE_S4, LFGC_SNAKE_S4), pack2(LFGC_SNAKE_S3, LFGC_SNAKE_S3));
    t = fma2(t, u, pack2(LFGC_SNAKE_S2, LFGC_SNAKE_S2));
    t = fma2(t, u, pack2(LFGC_SNAKE_S1, LFGC_SNAKE_S1));
    const unsigned long long ru = fma2(r, u, zero);
    t = fma2(t, ru, r);                                                                  // sin r
    float b0, b1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(b0), "=f"(b1) : "l"(biased));
    const int s0 = __float_as_int(b0) << 31, s1 = __float_as_int(b1) << 31;              // parity of n -> sign bit
    const unsigned long long sgnh = pack2(__int_as_float(0xBF000000 ^ s0), __int_as_float(0xBF000000 ^ s1));   // -0.5 (-1)^n
    const unsigned long long sgng = pack2(__int_as_float(0x3F800000 ^ s0), __int_as_float(0x3F800000 ^ s1));   // (-1)^n
    const unsigned long long h = fma2(sgnh, p, fma2(half, z, half));
    const unsigned long long g = fma2(sgng, t, half);
    asm("mov.b64 {%0, %1}, %2;" : "=f"(h0), "=f"(h1) : "l"(h));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(g0), "=f"(g1) : "l"(g));
}

// sin and cos of TWO arguments at once (the Fourier features of the tensor-core kernels): the reduction modulo pi and the
// two polynomials of snake_and_grad_fast2 applied to x itself; the sign (-1)^n goes onto both results with one XOR each.
// Max abs error 1.1e-7 (sin) / 6.2e-8 (cos) for |x| < 1e4; ~27 issue slots per pair against 2 x 28 for sincos_cw.
__device__ __forceinline__ void sincos_fast2(float x0, float x1, float& s0, float& c0, float& s1, float& c1) {
    const unsigned long long w = pack2(x0, x1);
    const unsigned long long one = pack2(1.0f, 1.0f), magic = pack2(12582912.0f, 12582912.0f), zero = pack2(0.0f, 0.0f);
    const unsigned long long biased = fma2(w, pack2(0.318309886f, 0.318309886f), magic);
    const unsigned long long n = fma2(biased, one, pack2(-12582912.0f, -12582912.0f));
    unsigned long long r = fma2(n, pack2(-3.14159274f, -3.14159274f), w);
    r = fma2(n, pack2(8.74227766e-08f, 8.74227766e-08f), r);
    const unsigned long long u = fma2(r, r, zero);
    unsigned long long p = fma2(u, pack2(LFGC_SNAKE_C5, LFGC_SNAKE_C5), pack2(LFGC_SNAKE_C4, LFGC_SNAKE_C4));
    p = fma2(p, u, pack2(LFGC_SNAKE_C3, LFGC_SNAKE_C3));
    p = fma2(p, u, pack2(LFGC_SNAKE_C2, LFGC_SNAKE_C2));
    p = fma2(p, u, pack2(-0.5f, -0.5f));
    p = fma2(p, u, one);                                                                 // cos r
    unsigned long long t = fma2(u, pack2(LFGC_SNAKE_S4, LFGC_SNAKE_S4), pack2(LFGC_SNAKE_S3, LFGC_SNAKE_S3));
    t = fma2(t, u, pack2(LFGC_SNAKE_S2, LFGC_SNAKE_S2));
    t = fma2(t, u, pack2(LFGC_SNAKE_S1, LFGC_SNAKE_S1));
    t = fma2(t, fma2(r, u, zero), r);                                                    // sin r
    float b0, b1, pc0, pc1, ps0, ps1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(b0), "=f"(b1) : "l"(biased));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(pc0), "=f"(pc1) : "l"(p));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(ps0), "=f"(ps1) : "l"(t));
    const int g0 = __float_as_int(b0) << 31, g1 = __float_as_int(b1) << 31;              // parity of n -> sign bit
    s0 = __int_as_float(__float_as_int(ps0) ^ g0);
    c0 = __int_as_float(__float_as_int(pc0) ^ g0);
    s1 = __int_as_float(__float_as_int(ps1) ^ g1);
    c1 = __int_as_float(__float_as_int(pc1) ^ g1);
}

// Hidden-layer activation of the fused kernels: ACT 0 = SnakeAlt (the fV-SRN decoder), ACT 1 = ReLU (Variance_Model,
// model/Variational_Dropout_Layer.py:159-175; gradient 1 where z > 0, as torch's threshold_backward).
template <int ACT>
__device__ __forceinline__ float act_value(float z) {
    if (ACT == 1) return fmaxf(z, 0.0f);
    return snake_precise(z);
}
template <int ACT>
__device__ __forceinline__ void act_value_grad(float z, float& h, float& g) {
    if (ACT == 1) {
        h = fmaxf(z, 0.0f);
        g = z > 0.0f ? 1.0f : 0.0f;
    } else {
        snake_and_grad_precise(z, h, g);
    }
}

// Philox4x32-10 counter-based generator (Salmon et al. 2011), one call per sample.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

// uniform voxel index in [0, n_voxels) from one Philox draw (multiply-high mapping of a 64-bit random)
__device__ __forceinline__ unsigned long long philox_voxel(uint64_t seed, uint64_t counter,
                                                           unsigned long long n_voxels) {
    uint4 r = philox4x32_10(make_uint4((uint32_t)counter, (uint32_t)(counter >> 32), 0x6c666763u, 0u),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    unsigned long long r64 = ((unsigned long long)r.x << 32) | r.y;
    return __umul64hi(r64, n_voxels);
}

// Normalised coordinate of voxel index i along an axis with max_idx = R-1 and scale = max_idx / max(max_idx):
// scales * ((1 - -1) * ((raw - 0) / (max_idx - 0)) + -1), fp32, one rounding per operation as in
// data/IndexDataset.py:7-8,92-95 (no FMA contraction).
__device__ __forceinline__ float normalized_coord(float raw, float max_idx, float scale) {
    float q = __fdiv_rn(raw, max_idx);
    float v = __fadd_rn(__fmul_rn(2.0f, q), -1.0f);
    return __fmul_rn(scale, v);
}

#endif  // __CUDACC__

}  // namespace lfgc
