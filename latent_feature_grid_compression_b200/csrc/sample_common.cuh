// Pieces shared by the forward and backward sample kernels: the per-sample input stage (trilinear gather of the
// channels-last latent grid + Fourier features) and the shared-memory MLP parameter staging.
#pragma once
#include "lfgc_common.cuh"

namespace lfgc {

constexpr int kTile = 128;     // samples per CTA tile (one per thread in the per-sample phases)
constexpr int kThreads = 128;  // 4 warps; warp w owns tile columns [32w, 32w+32)
constexpr int kMaxFreq = 16;

struct SampleParams {
    int C, Cp, H, L, F, in0, in0p;  // in0 = 3 + 6F + C real inputs; in0p = 3 + 6F + Cp rows of the input block
    int G[3];                       // (D, H, W) = (z, y, x)
    float omega[kMaxFreq];          // fp32(2^k) * 2 * fp32(pi), as model/Feature_Embedding.py:28-29 builds them
    int flags;
};

// The 8 trilinear corners of F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=False)
// (model/Feature_Grid_Model.py:63-64; ATen grid_sampler_unnormalize: ((coord+1)*size-1)/2, corner weights
// (x1-ix)(y1-iy)(z1-iz) ...).  coords[0] indexes W (x, last grid dim), [1] -> H, [2] -> D.
// Out-of-range corners get weight 0 and a clamped (valid) offset.
struct Corners {
    float w[8];
    int off[8];  // element offset of the corner's channel vector in the channels-last grid
};

__device__ __forceinline__ void axis_corners(float coord, int size, int& i0, int& i1, float& w0, float& w1) {
    const float f = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(coord, 1.0f), (float)size), 1.0f), 2.0f);
    const float fl = floorf(f);
    w0 = __fsub_rn(__fadd_rn(fl, 1.0f), f);
    w1 = __fsub_rn(f, fl);
    // guard the float->int conversion against huge/NaN coordinates
    const float flc = fminf(fmaxf(fl, -2.0f), (float)size + 1.0f);
    const int i = (int)flc;
    const bool finite_ok = (fl == flc);
    i0 = i;
    i1 = i + 1;
    if (!finite_ok || i0 < 0 || i0 >= size) { w0 = 0.0f; i0 = 0; }
    if (!finite_ok || i1 < 0 || i1 >= size) { w1 = 0.0f; i1 = 0; }
}

__device__ __forceinline__ void make_corners(const SampleParams& P, float cx, float cy, float cz, Corners& K) {
    int x0, x1, y0, y1, z0, z1;
    float wx0, wx1, wy0, wy1, wz0, wz1;
    axis_corners(cx, P.G[2], x0, x1, wx0, wx1);
    axis_corners(cy, P.G[1], y0, y1, wy0, wy1);
    axis_corners(cz, P.G[0], z0, z1, wz0, wz1);
    const int xs[2] = {x0, x1}, ys[2] = {y0, y1}, zs[2] = {z0, z1};
    const float wxs[2] = {wx0, wx1}, wys[2] = {wy0, wy1}, wzs[2] = {wz0, wz1};
#pragma unroll
    for (int c = 0; c < 8; ++c) {  // ATen order: tnw, tne, tsw, tse, bnw, bne, bsw, bse
        const int dz = c >> 2, dy = (c >> 1) & 1, dx = c & 1;
        K.w[c] = __fmul_rn(__fmul_rn(wxs[dx], wys[dy]), wzs[dz]);
        K.off[c] = ((zs[dz] * P.G[1] + ys[dy]) * P.G[2] + xs[dx]) * P.Cp;
    }
}

// Writes the input block of one sample into column `col` of X (row stride S):
// rows [0,3) xyz | [3, 3+6F) Fourier | [3+6F, 3+6F+Cp) gathered features (pad channels are zero in the grid).
__device__ __forceinline__ void stage_inputs(const SampleParams& P, const float* __restrict__ grid, float cx,
                                             float cy, float cz, const Corners& K, float* X, int S, int col) {
    X[0 * S + col] = cx;
    X[1 * S + col] = cy;
    X[2 * S + col] = cz;
    for (int f = 0; f < P.F; ++f) {
        const float om = P.omega[f];
        float s, c;
        sincos_cw(__fmul_rn(cx, om), s, c);  // the argument is rounded to fp32 first (Feature_Embedding.py:33)
        X[(3 + 6 * f + 0) * S + col] = s;
        X[(3 + 6 * f + 3) * S + col] = c;
        sincos_cw(__fmul_rn(cy, om), s, c);
        X[(3 + 6 * f + 1) * S + col] = s;
        X[(3 + 6 * f + 4) * S + col] = c;
        sincos_cw(__fmul_rn(cz, om), s, c);
        X[(3 + 6 * f + 2) * S + col] = s;
        X[(3 + 6 * f + 5) * S + col] = c;
    }
    const int fbase = 3 + 6 * P.F;
    for (int c4 = 0; c4 < P.Cp; c4 += 4) {
        float4 v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = ldg_f4(grid + K.off[c] + c4);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            a.x = fmaf(v[c].x, K.w[c], a.x);
            a.y = fmaf(v[c].y, K.w[c], a.y);
            a.z = fmaf(v[c].z, K.w[c], a.z);
            a.w = fmaf(v[c].w, K.w[c], a.w);
        }
        X[(fbase + c4 + 0) * S + col] = a.x;
        X[(fbase + c4 + 1) * S + col] = a.y;
        X[(fbase + c4 + 2) * S + col] = a.z;
        X[(fbase + c4 + 3) * S + col] = a.w;
    }
}

// Forward-layout weights in shared memory: for layer l a [K_l][HP] block (k-major, outputs contiguous) so that
// a lane reads its NO consecutive output weights of one k with 16-byte loads.  Offsets in floats.
template <int HP>
struct FwdWeights {
    __device__ __forceinline__ static int layer_off(int l, int in0) { return l == 0 ? 0 : in0 * HP + (l - 1) * HP * HP; }
    __device__ __forceinline__ static int total(int L, int in0) { return in0 * HP + (L - 1) * HP * HP; }
};

// global packed layout: [W0 (H x in0) | b0 (H) | W1 (H x H) | b1 (H) | ... | Wf (H) | bf]
__device__ __forceinline__ int mlp_w_off(int l, int in0, int H) {
    return l == 0 ? 0 : (in0 * H + H) + (l - 1) * (H * H + H);
}
__device__ __forceinline__ int mlp_b_off(int l, int in0, int H) { return mlp_w_off(l, in0, H) + (l == 0 ? in0 : H) * H; }
__device__ __forceinline__ int mlp_wf_off(int L, int in0, int H) { return mlp_w_off(L, in0, H); }

// Stage the MLP parameters into shared memory.  Global reads run along the packed buffer (coalesced; 16-byte
// vectors where the layer's block is aligned), the transposition happens on the shared-memory side.
// SRC_SMEM: `mlp` points at a copy of the packed block in shared memory (plain loads) instead of global memory (__ldg).
template <bool SRC_SMEM>
__device__ __forceinline__ float ld_param(const float* p) { return SRC_SMEM ? *p : __ldg(p); }
template <bool SRC_SMEM>
__device__ __forceinline__ float4 ld_param4(const float4* p) { return SRC_SMEM ? *p : __ldg(p); }

template <int HP, bool SRC_SMEM = false>
__device__ __forceinline__ void load_fwd_weights(const SampleParams& P, const float* __restrict__ mlp, float* Wt,
                                                 float* bias, float* Wf, float* bf) {
    const int H = P.H, in0 = P.in0;
    const int total = FwdWeights<HP>::total(P.L, in0);
    if (H < HP) {  // pad columns must read as zero
        for (int e = threadIdx.x; e < total; e += blockDim.x) Wt[e] = 0.0f;
        __syncthreads();
    }
    for (int l = 0; l < P.L; ++l) {
        const int K = l == 0 ? in0 : H;
        const int woff = mlp_w_off(l, in0, H);
        const float* W = mlp + woff;
        float* dst = Wt + FwdWeights<HP>::layer_off(l, in0);
        if ((K & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
            const int n4 = (H * K) >> 2;
            for (int e4 = threadIdx.x; e4 < n4; e4 += blockDim.x) {
                const float4 v = ld_param4<SRC_SMEM>(reinterpret_cast<const float4*>(W) + e4);
                const int e = e4 << 2;
                const int j = e / K, k = e - j * K;
                dst[(k + 0) * HP + j] = v.x;
                dst[(k + 1) * HP + j] = v.y;
                dst[(k + 2) * HP + j] = v.z;
                dst[(k + 3) * HP + j] = v.w;
            }
        } else {
            for (int e = threadIdx.x; e < H * K; e += blockDim.x) {
                const int j = e / K, k = e - j * K;
                dst[k * HP + j] = ld_param<SRC_SMEM>(W + e);
            }
        }
        const float* b = mlp + mlp_b_off(l, in0, H);
        for (int j = threadIdx.x; j < HP; j += blockDim.x) bias[l * HP + j] = j < H ? ld_param<SRC_SMEM>(b + j) : 0.0f;
    }
    const float* wf = mlp + mlp_wf_off(P.L, in0, H);
    for (int j = threadIdx.x; j < HP; j += blockDim.x) Wf[j] = j < H ? ld_param<SRC_SMEM>(wf + j) : 0.0f;
    if (threadIdx.x == 0) *bf = ld_param<SRC_SMEM>(wf + H);
}

// Backward-layout weights: Wb_l[j][k] = W_l[j][k] (l >= 1, row stride HP) and the layer-0 feature columns
// W0f[j][c] = W_0[j][3+6F+c] (row stride Cp); pads are zero.
template <int HP, bool SRC_SMEM = false>
__device__ __forceinline__ void load_bwd_weights(const SampleParams& P, const float* __restrict__ mlp, float* Wb,
                                                 float* W0f) {
    const int H = P.H, in0 = P.in0;
    for (int l = 1; l < P.L; ++l) {
        const int woff = mlp_w_off(l, in0, H);
        const float* W = mlp + woff;
        float* dst = Wb + (l - 1) * HP * HP;
        if (H == HP && (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
            for (int e4 = threadIdx.x; e4 < (HP * HP) >> 2; e4 += blockDim.x)
                reinterpret_cast<float4*>(dst)[e4] = ld_param4<SRC_SMEM>(reinterpret_cast<const float4*>(W) + e4);
        } else {
            for (int e = threadIdx.x; e < HP * HP; e += blockDim.x) {
                const int j = e / HP, k = e % HP;
                dst[e] = (j < H && k < H) ? ld_param<SRC_SMEM>(W + j * H + k) : 0.0f;
            }
        }
    }
    for (int e = threadIdx.x; e < HP * P.Cp; e += blockDim.x) {
        const int j = e / P.Cp, c = e % P.Cp;
        W0f[e] = (j < H && c < P.C) ? ld_param<SRC_SMEM>(mlp + j * in0 + 3 + 6 * P.F + c) : 0.0f;
    }
}

// Packed fp32 FMA (sm_100+): d.xy += a * b.xy.  One issue slot for two FMAs; measured on B200: 117 FMA/clk/SM
// with FFMA2 against 64 FMA/clk/SM with scalar 3-register FFMA (profiles/microbench/fma_bench.cu), so every contraction uses it.
__device__ __forceinline__ void ffma2(float2& d, float a, float2 b) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ra) : "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rd) : "f"(d.x), "f"(d.y));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rd) : "l"(ra), "l"(rb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
}
__device__ __forceinline__ void ffma2v(float2& d, float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rd) : "f"(d.x), "f"(d.y));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rd) : "l"(ra), "l"(rb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
}

// One dense layer for a warp's 32 samples, register-tiled 4 samples x NO outputs per lane:
// acc[s][o] = bias[o] + sum_k in[k][col0+s] * Wt[k][j0+o].  `in` has row stride S.  Packed FFMA2 over output pairs.
template <int HP, int S>
__device__ __forceinline__ void warp_gemm(const float* __restrict__ in, const float* __restrict__ Wt,
                                          const float* __restrict__ bias, int K, int col0, int j0,
                                          float (&acc)[4][HP / 4]) {
    constexpr int NO = HP / 4;
    float2 a2[4][NO / 2];
#pragma unroll
    for (int o = 0; o < NO; o += 2) {
        const float b0 = bias ? bias[j0 + o] : 0.0f, b1 = bias ? bias[j0 + o + 1] : 0.0f;
#pragma unroll
        for (int s = 0; s < 4; ++s) a2[s][o >> 1] = make_float2(b0, b1);
    }
    const float* inp = in + col0;
    const float* wp = Wt + j0;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(inp + k * S);
#pragma unroll
        for (int o4 = 0; o4 < NO; o4 += 4) {
            const float4 t = *reinterpret_cast<const float4*>(wp + k * HP + o4);
            const float2 w01 = make_float2(t.x, t.y), w23 = make_float2(t.z, t.w);
            ffma2(a2[0][o4 >> 1], a.x, w01); ffma2(a2[0][(o4 >> 1) + 1], a.x, w23);
            ffma2(a2[1][o4 >> 1], a.y, w01); ffma2(a2[1][(o4 >> 1) + 1], a.y, w23);
            ffma2(a2[2][o4 >> 1], a.z, w01); ffma2(a2[2][(o4 >> 1) + 1], a.z, w23);
            ffma2(a2[3][o4 >> 1], a.w, w01); ffma2(a2[3][(o4 >> 1) + 1], a.w, w23);
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int o = 0; o < NO; o += 2) {
            acc[s][o] = a2[s][o >> 1].x;
            acc[s][o + 1] = a2[s][o >> 1].y;
        }
}

// internal flag (not part of the ABI): ReLU hidden activations, no grid / embedding inputs (Variance_Model)
constexpr int kFlagPlainRelu = 1 << 16;

// Parameters of the plain coordinate MLP 3 -> H x L (ReLU) -> 1 (model/Variational_Dropout_Layer.py:159-175): the
// fused kernels run it as a model without grid features and Fourier rows.
inline int fill_plain_params(int H, int L, SampleParams& P) {
    if (H < 1 || H > 32) return fail(LFGC_E_UNSUPPORTED, "plain MLP: size_layers=%d unsupported (1..32)", H);
    if (L < 1 || L > 4) return fail(LFGC_E_UNSUPPORTED, "plain MLP: n_layers=%d unsupported (1..4)", L);
    P.C = 0; P.Cp = 0; P.H = H; P.L = L; P.F = 0;
    P.in0 = 3; P.in0p = 3;
    for (int a = 0; a < 3; ++a) P.G[a] = 1;
    for (int f = 0; f < kMaxFreq; ++f) P.omega[f] = 0.0f;
    P.flags = kFlagPlainRelu;
    return LFGC_OK;
}

inline int fill_sample_params(const lfgc_model_desc* m, int flags, SampleParams& P) {
    if (!m) return fail(LFGC_E_INVALID, "model desc is null");
    if (m->C < 1 || m->Cp < m->C || (m->Cp & 3)) return fail(LFGC_E_INVALID, "bad C=%d Cp=%d", m->C, m->Cp);
    if (m->H < 1 || m->H > 64) return fail(LFGC_E_UNSUPPORTED, "n_hidden_size=%d unsupported (1..64)", m->H);
    if (m->L < 1 || m->L > LFGC_MAX_LAYERS) return fail(LFGC_E_UNSUPPORTED, "n_layers=%d unsupported (1..%d)", m->L, LFGC_MAX_LAYERS);
    if (m->F < 0 || m->F > kMaxFreq) return fail(LFGC_E_UNSUPPORTED, "n_embedding_freq=%d unsupported (0..%d)", m->F, kMaxFreq);
    for (int a = 0; a < 3; ++a)
        if (m->G[a] < 1) return fail(LFGC_E_INVALID, "grid extent %d", m->G[a]);
    if ((int64_t)m->G[0] * m->G[1] * m->G[2] * m->Cp >= (int64_t)1 << 31)
        return fail(LFGC_E_UNSUPPORTED, "grid too large for 32-bit offsets");
    P.C = m->C; P.Cp = m->Cp; P.H = m->H; P.L = m->L; P.F = m->F;
    P.in0 = 3 + 6 * m->F + m->C;
    P.in0p = 3 + 6 * m->F + m->Cp;
    for (int a = 0; a < 3; ++a) P.G[a] = m->G[a];
    for (int f = 0; f < kMaxFreq; ++f) {
        // 2.**linspace(0, F-1, F) * 2. * np.pi in fp32 (model/Feature_Embedding.py:28-29)
        float fb = ldexpf(1.0f, f);
        fb = fb * 2.0f;
        fb = fb * (float)3.141592653589793;
        P.omega[f] = fb;
    }
    P.flags = flags;
    return LFGC_OK;
}

}  // namespace lfgc
