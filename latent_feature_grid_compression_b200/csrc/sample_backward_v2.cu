// Fused training kernel, wide variant (the one the shipped configurations run): same mathematics as the generic
// kernel in sample_backward.cu, re-organised for issue-slot and latency efficiency on sm_100a:
//
//  * 4 to 8 warps per CTA (one CTA per SM), 16 samples per warp: 2 resident warps per scheduler instead of 1 (wider
//    CTAs are register-limited to 168 / 128 registers per thread and spill: measured slower),
//    which is what hides the shared-memory / L2 latencies the generic kernel exposes (ncu: issue-active 32 %).
//  * lane = 4 samples x 4 outputs; the multiply-accumulates are packed FFMA2 (fma.rn.f32x2, sm_100+): one issue
//    slot per two FMAs, the activation broadcast folded into the instruction's scalar operand.
//  * S'(z_l) never touches shared memory: the lane that produces z_l[s][o] in the forward is the lane that needs
//    S'(z_l[s][o]) in the backward, so it stays in registers (16 per layer).  Shared memory per sample drops from
//    1092 B to ~770 B, which is what makes the wider tile fit.
//  * dW_l = dz_l^T h_{l-1} is evaluated right after dz_l is produced (one CTA barrier per layer), by warp
//    (row group, sample group) with the partial sums kept in registers across the whole persistent loop; they leave
//    the SM once, as one plain store per entry into a per-(CTA, sample group) slice of the workspace.
//  * the per-sample phases (gather + embedding, scatter) use two lanes per sample and stay warp-local.
#include "sample_backward.cuh"

#include <stdlib.h>

// Debug-only phase timing: build with -DLFGC_PHASE_TIMING and read lfgc_phase_timing() (profiles/phase_timing.py).
#ifdef LFGC_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[16];
#define PHASE_DECL long long _pt_last = clock64(); unsigned long long _pt[12] = {0,0,0,0,0,0,0,0,0,0,0,0};
#define PHASE_MARK(i) { long long _n = clock64(); _pt[i] += (unsigned long long)(_n - _pt_last); _pt_last = _n; }
#define PHASE_FLUSH() { if ((threadIdx.x & 31) == 0) for (int _i = 0; _i < 12; ++_i) atomicAdd(&g_phase_cycles[_i], _pt[_i]); }
#else
#define PHASE_DECL
#define PHASE_MARK(i)
#define PHASE_FLUSH()
#endif

namespace lfgc {
namespace v2 {

constexpr int HP = 32;
constexpr int LMAX = 4;

constexpr int kBaccStride = LMAX * HP + HP + 4;

struct Layout {
    int Wt, bias, Wf, bf, Wb, W0f, bacc, X, H, DZ, DF, AUX, total, S;
};

__host__ __device__ inline int x_rows(const SampleParams& P) { return (P.in0p + 31) & ~31; }

__host__ __device__ inline Layout make_layout(const SampleParams& P, int nw) {
    Layout o;
    o.S = 16 * nw + 4;
    int p = 0;
    o.Wt = p;   p += P.in0 * HP + (P.L - 1) * HP * HP;
    o.bias = p; p += P.L * HP;
    o.Wf = p;   p += HP;
    o.bf = p;   p += 4;
    o.Wb = p;   p += (P.L - 1) * HP * HP;
    o.W0f = p;  p += HP * P.Cp;
    o.bacc = p; p += nw * kBaccStride;      // per warp: bias gradients | Wf gradient | bf gradient | loss
    o.X = p;    p += x_rows(P) * o.S;     // padded so that the weight-gradient phase needs no column clamp
    o.H = p;    p += (P.L - 1) * HP * o.S;
    o.DZ = p;   p += 2 * HP * o.S;
    if (P.L >= 2 && P.Cp <= HP) {
        o.DF = o.H + (P.L - 2) * HP * o.S;  // d(features) re-uses the rows of h_{L-2} (dead after dW_{L-1})
    } else {
        o.DF = p;
        p += P.Cp * o.S;
    }
    o.AUX = p;  p += 5 * o.S;              // 0: ground truth | grad_out, 1: valid, 2..4: coordinates
    o.total = p;
    return o;
}

// acc[s][p] = (out 2p, out 2p+1) of sample s:  acc = bias + sum_k in[k][s] * W[k][0..3]
// `in` points at the lane's 4 samples of row 0, `W` at the lane's 4 outputs of row 0; S and LDW are compile-time so
// that every load of the unrolled body is base + immediate.
template <int S, int LDW>
__device__ __forceinline__ void gemm16(const float* __restrict__ in, const float* __restrict__ W,
                                       const float* __restrict__ bias4, int K, float2 (&acc)[4][2]) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias4) b = *reinterpret_cast<const float4*>(bias4);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        acc[s][0] = make_float2(b.x, b.y);
        acc[s][1] = make_float2(b.z, b.w);
    }
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(in + k * S);
        const float4 w = *reinterpret_cast<const float4*>(W + k * LDW);
        const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
        ffma2(acc[0][0], a.x, w01); ffma2(acc[0][1], a.x, w23);
        ffma2(acc[1][0], a.y, w01); ffma2(acc[1][1], a.y, w23);
        ffma2(acc[2][0], a.z, w01); ffma2(acc[2][1], a.z, w23);
        ffma2(acc[3][0], a.w, w01); ffma2(acc[3][1], a.w, w23);
    }
}
// same with a run-time weight row stride (the layer-0 feature columns)
template <int S>
__device__ __forceinline__ void gemm16_ldw(const float* __restrict__ in, const float* __restrict__ W, int ldw, int K,
                                           float2 (&acc)[4][2]) {
#pragma unroll
    for (int s = 0; s < 4; ++s) acc[s][0] = acc[s][1] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(in + k * S);
        const float4 w = *reinterpret_cast<const float4*>(W + k * ldw);
        const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
        ffma2(acc[0][0], a.x, w01); ffma2(acc[0][1], a.x, w23);
        ffma2(acc[1][0], a.y, w01); ffma2(acc[1][1], a.y, w23);
        ffma2(acc[2][0], a.z, w01); ffma2(acc[2][1], a.z, w23);
        ffma2(acc[3][0], a.w, w01); ffma2(acc[3][1], a.w, w23);
    }
}

__device__ __forceinline__ float pick(const float2 (&acc)[4][2], int s, int o) {
    return (o & 1) ? acc[s][o >> 1].y : acc[s][o >> 1].x;
}

// input stage for one sample shared by two lanes: half 0 writes xyz + Fourier rows and the even feature vectors,
// half 1 the odd feature vectors
__device__ __forceinline__ void stage_inputs_half(const SampleParams& P, const float* __restrict__ grid, float cx,
                                                  float cy, float cz, const Corners& K, float* X, int S, int col,
                                                  int half) {
    if (half == 0) {
        X[0 * S + col] = cx;
        X[1 * S + col] = cy;
        X[2 * S + col] = cz;
        for (int f = 0; f < P.F; ++f) {
            const float om = P.omega[f];
            float s, c;
            sincos_cw(__fmul_rn(cx, om), s, c);
            X[(3 + 6 * f + 0) * S + col] = s;
            X[(3 + 6 * f + 3) * S + col] = c;
            sincos_cw(__fmul_rn(cy, om), s, c);
            X[(3 + 6 * f + 1) * S + col] = s;
            X[(3 + 6 * f + 4) * S + col] = c;
            sincos_cw(__fmul_rn(cz, om), s, c);
            X[(3 + 6 * f + 2) * S + col] = s;
            X[(3 + 6 * f + 5) * S + col] = c;
        }
    }
    const int fbase = 3 + 6 * P.F;
    for (int c4 = 4 * half; c4 < P.Cp; c4 += 8) {
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

// dW partial of one (2 rows x 4 columns) block over this warp's sample blocks.  dzp = row `row0` of dz_l at the
// warp's first sample block, hp = column `cb` of h_{l-1} likewise; the second row is 4 rows further, the columns
// are 8 rows apart: everything is base + immediate.  step = floats between this warp's consecutive sample blocks.
template <int S, int TILE>
__device__ __forceinline__ void dw_block(const float* __restrict__ dzp, const float* __restrict__ hp, int n0, int step,
                                         float (&out)[2][4]) {
    float2 t[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) t[r][c] = make_float2(0.f, 0.f);
    dzp += n0;
    hp += n0;
    // software pipeline over two register buffers (A/B): the loads of sample block i+1 are in flight while block i
    // is multiplied, without register copies
    float4 dzA[2], hvA[4], dzB[2], hvB[4];
    auto load = [&](float4 (&dz)[2], float4 (&hv)[4], int off) {
#pragma unroll
        for (int r = 0; r < 2; ++r) dz[r] = *reinterpret_cast<const float4*>(dzp + off + r * 4 * S);
#pragma unroll
        for (int c = 0; c < 4; ++c) hv[c] = *reinterpret_cast<const float4*>(hp + off + c * 8 * S);
    };
    auto mac = [&](const float4 (&dz)[2], const float4 (&hv)[4]) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                ffma2v(t[r][c], make_float2(dz[r].x, dz[r].y), make_float2(hv[c].x, hv[c].y));
                ffma2v(t[r][c], make_float2(dz[r].z, dz[r].w), make_float2(hv[c].z, hv[c].w));
            }
    };
    load(dzA, hvA, 0);
    while (true) {
        if (n0 + step < TILE) load(dzB, hvB, step);
        mac(dzA, hvA);
        if (n0 + step >= TILE) break;
        if (n0 + 2 * step < TILE) load(dzA, hvA, 2 * step);
        mac(dzB, hvB);
        if (n0 + 2 * step >= TILE) break;
        n0 += 2 * step;
        dzp += 2 * step;
        hp += 2 * step;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) out[r][c] += t[r][c].x + t[r][c].y;
}

// NW (warps per CTA) is a template parameter so that the tile size and the row stride are compile-time constants; it
// also bounds the registers per thread (registers are per scheduler: 16384 / warps per scheduler / 32):
// 8 warps -> 255, 12 -> 168, 16 -> 128.
template <int FUSED, int NC0, int NW, int ACT>
__global__ void __launch_bounds__(32 * NW, 1) backward_v2_kernel(const __grid_constant__ BwdArgs A) {
    LFGC_PDL_PROLOGUE();
    extern __shared__ __align__(16) float smem[];
    PHASE_DECL
    const SampleParams& P = A.P;
    constexpr int TILE = 16 * NW;
    constexpr int S = TILE + 4;
    constexpr int KGmax = (NW + 3) >> 2;      // slices per CTA
    const Layout Lo = make_layout(P, NW);
    float* Wt = smem + Lo.Wt;
    float* bias = smem + Lo.bias;
    float* Wf = smem + Lo.Wf;
    float* Wb = smem + Lo.Wb;
    float* W0f = smem + Lo.W0f;
    float* bacc = smem + Lo.bacc;
    float* X = smem + Lo.X;
    float* Hs = smem + Lo.H;
    float* DZ = smem + Lo.DZ;
    float* DF = smem + Lo.DF;
    float* AUX = smem + Lo.AUX;

    const int H = P.H, in0 = P.in0, L = P.L;
    // Parameter staging: one pass of independent 16-byte loads brings the packed block into shared memory (a single L2
    // round trip instead of one per loop iteration of the transposing loaders), the layouts are then built from there.
    // The copy lives in the dz double buffer, which is dead until the first tile's backward pass.
    if ((A.pcount + 3) / 4 * 4 <= 2 * HP * S && (reinterpret_cast<uintptr_t>(A.mlp) & 15) == 0) {
        float* stage = DZ;
        const int n4 = A.pcount >> 2;
        constexpr int NT = 32 * NW;
        float4 v[6];
#pragma unroll
        for (int i = 0; i < 6; ++i)
            if ((int)threadIdx.x + i * NT < n4) v[i] = __ldg(reinterpret_cast<const float4*>(A.mlp) + threadIdx.x + i * NT);
#pragma unroll
        for (int i = 0; i < 6; ++i)
            if ((int)threadIdx.x + i * NT < n4) reinterpret_cast<float4*>(stage)[threadIdx.x + i * NT] = v[i];
        for (int e4 = threadIdx.x + 6 * NT; e4 < n4; e4 += NT)
            reinterpret_cast<float4*>(stage)[e4] = __ldg(reinterpret_cast<const float4*>(A.mlp) + e4);
        if ((int)threadIdx.x < (A.pcount & 3)) stage[4 * n4 + threadIdx.x] = __ldg(A.mlp + 4 * n4 + threadIdx.x);
        __syncthreads();
        load_fwd_weights<HP, true>(P, stage, Wt, bias, Wf, smem + Lo.bf);
        load_bwd_weights<HP, true>(P, stage, Wb, W0f);
    } else {
        load_fwd_weights<HP>(P, A.mlp, Wt, bias, Wf, smem + Lo.bf);
        load_bwd_weights<HP>(P, A.mlp, Wb, W0f);
    }
    for (int e = threadIdx.x; e < NW * kBaccStride; e += blockDim.x) bacc[e] = 0.0f;
    for (int e = P.in0p * S + threadIdx.x; e < x_rows(P) * S; e += blockDim.x) X[e] = 0.0f;  // pad rows of the input block
    __syncthreads();
    const float bf = smem[Lo.bf];
    PHASE_MARK(0)  // setup

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // sample-major GEMMs: lane = 4 samples x 4 outputs
    const int og = lane >> 2, sg = lane & 3;
    const int col0 = 16 * warp + 4 * sg;
    const int j0 = 4 * og;
    // per-sample phases: two lanes per sample
    const int scol = 16 * warp + (lane >> 1);
    const int half = lane & 1;
    // weight-gradient phase: warp = (row group, sample group), lane = 2 rows x 4 columns
    float* wacc = bacc + warp * kBaccStride;  // this warp's private accumulators
    const int rg = warp & 3, kg = warp >> 2;
    const int KG = (NW - rg + 3) >> 2;        // warps sharing this row group: they split the tile's sample blocks
    const int row0 = 8 * rg + (lane >> 3);
    const int cb = lane & 7;

    float dWacc[LMAX][2][4];
    float dW0x[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            dW0x[r][c] = 0.0f;
#pragma unroll
            for (int l = 0; l < LMAX; ++l) dWacc[l][r][c] = 0.0f;
        }
    float accWf[4] = {0.f, 0.f, 0.f, 0.f};
    float accbf = 0.0f, loss_part = 0.0f;

    uint64_t sample_base = A.sample_offset;
    if (FUSED && A.step_dev) sample_base += (uint64_t)(*A.step_dev) * A.step_stride;

    const int64_t ntiles = (A.n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- G: input stage, two lanes per sample (warp-local) -----------------------------------------------------
        {
            const int64_t s = tile * TILE + scol;
            const bool valid = s < A.n;
            float cx = 0.f, cy = 0.f, cz = 0.f, aux = 0.f;
            if (valid) {
                if (FUSED && A.coords) {  // host-fed step: caller-supplied positions and target values
                    cx = __ldg(A.coords + 3 * s);
                    cy = __ldg(A.coords + 3 * s + 1);
                    cz = __ldg(A.coords + 3 * s + 2);
                    if (half == 0) aux = __ldg(A.grad_out + s);
                } else if (FUSED) {
                    unsigned long long v = A.explicit_idx ? (unsigned long long)A.explicit_idx[s]
                                                          : philox_voxel(A.seed, sample_base + (uint64_t)s, A.n_voxels);
                    const unsigned long long r12 = (unsigned long long)A.R[1] * A.R[2];
                    const int i = (int)(v / r12);
                    const int j = (int)((v / A.R[2]) % A.R[1]);
                    const int k = (int)(v % A.R[2]);
                    cx = normalized_coord((float)i, A.max_idx[0], A.scales[0]);
                    cy = normalized_coord((float)j, A.max_idx[1], A.scales[1]);
                    cz = normalized_coord((float)k, A.max_idx[2], A.scales[2]);
                    if (half == 0) aux = __ldg(A.volume + v);
                } else {
                    cx = __ldg(A.coords + 3 * s);
                    cy = __ldg(A.coords + 3 * s + 1);
                    cz = __ldg(A.coords + 3 * s + 2);
                    if (half == 0) aux = __ldg(A.grad_out + s);
                }
            }
            Corners K;
            make_corners(P, cx, cy, cz, K);
            stage_inputs_half(P, A.grid, cx, cy, cz, K, X, S, scol, half);
            if (half == 0) {
                AUX[scol] = aux;
                AUX[S + scol] = valid ? 1.0f : 0.0f;
                AUX[2 * S + scol] = cx;
                AUX[3 * S + scol] = cy;
                AUX[4 * S + scol] = cz;
            }
        }
        __syncwarp();
        PHASE_MARK(1)  // input stage

        // ---- F: forward; S'(z) stays in registers -----------------------------------------------------------------------
        float greg[LMAX][16];
        float hl[4][4];
#pragma unroll
        for (int l = 0; l < LMAX; ++l) {
            if (l < L) {
                float2 acc[4][2];
                const float* in = l == 0 ? X : Hs + (l - 1) * HP * S;
                gemm16<S, HP>(in + col0, Wt + FwdWeights<HP>::layer_off(l, in0) + j0, bias + l * HP + j0,
                              l == 0 ? in0 : HP, acc);
                float* hrow = Hs + l * HP * S;
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    float h[4];
#pragma unroll
                    for (int s = 0; s < 4; ++s) act_value_grad<ACT>(pick(acc, s, o), h[s], greg[l][s * 4 + o]);
                    if (l + 1 < L) {
                        *reinterpret_cast<float4*>(hrow + (j0 + o) * S + col0) = make_float4(h[0], h[1], h[2], h[3]);
                    } else {
#pragma unroll
                        for (int s = 0; s < 4; ++s) hl[s][o] = h[s];
                    }
                }
                __syncwarp();
            }
        }
        PHASE_MARK(2)  // forward layers
        float y[4] = {0.f, 0.f, 0.f, 0.f};
        const float4 wf4 = *reinterpret_cast<const float4*>(Wf + j0);
        const float wf[4] = {wf4.x, wf4.y, wf4.z, wf4.w};
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
            for (int s = 0; s < 4; ++s) y[s] = fmaf(hl[s][o], wf[o], y[s]);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            y[s] += __shfl_xor_sync(0xffffffffu, y[s], 4);
            y[s] += __shfl_xor_sync(0xffffffffu, y[s], 8);
            y[s] += __shfl_xor_sync(0xffffffffu, y[s], 16);
        }
        float dy[4];
        {
            const float4 a4 = *reinterpret_cast<const float4*>(AUX + col0);
            const float4 v4 = *reinterpret_cast<const float4*>(AUX + S + col0);
            const float aux[4] = {a4.x, a4.y, a4.z, a4.w};
            const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (FUSED) {
                    const float e = (y[s] + bf) - aux[s];
                    dy[s] = vv[s] * A.loss_scale2 * e;
                    if (og == 0) loss_part = fmaf(vv[s] * e, e, loss_part);
                    if (A.log_sigma && vv[s] != 0.0f) {   // Gaussian likelihood with per-sample log sigma
                        const int64_t gs = tile * TILE + col0 + s;
                        const float w = expf(-2.0f * __ldg(A.log_sigma + gs));
                        dy[s] *= w;
                        if (A.dlog_sigma && og == 0) A.dlog_sigma[gs] = A.loss_scale2 * (1.0f - e * e * w);
                    }
                } else {
                    dy[s] = aux[s];
                }
            }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
            for (int s = 0; s < 4; ++s) accWf[o] = fmaf(dy[s], hl[s][o], accWf[o]);
        if (og == 0) accbf += (dy[0] + dy[1]) + (dy[2] + dy[3]);

        // ---- B: backward, one CTA barrier per layer ------------------------------------------------------------------------
        float2 dh[4][2];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            dh[s][0] = make_float2(dy[s] * wf[0], dy[s] * wf[1]);
            dh[s][1] = make_float2(dy[s] * wf[2], dy[s] * wf[3]);
        }
#pragma unroll
        for (int l = LMAX - 1; l >= 0; --l) {
            if (l < L) {
                float* dzb = DZ + ((L - 1 - l) & 1) * HP * S;
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    float4 dz;
                    dz.x = pick(dh, 0, o) * greg[l][0 + o];
                    dz.y = pick(dh, 1, o) * greg[l][4 + o];
                    dz.z = pick(dh, 2, o) * greg[l][8 + o];
                    dz.w = pick(dh, 3, o) * greg[l][12 + o];
                    *reinterpret_cast<float4*>(dzb + (j0 + o) * S + col0) = dz;
                    float v = (dz.x + dz.y) + (dz.z + dz.w);
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if (sg == 0) wacc[l * HP + j0 + o] += v;  // owned by this lane: no atomics
                }
                PHASE_MARK(3)  // dz + bias
                __syncthreads();  // dz_l of every warp is visible; every warp has finished layer l+1
                PHASE_MARK(4)  // barrier wait

                // weight gradient of layer l: contraction over this warp's share of the tile's samples
                {
                    const float* hin = l == 0 ? X : Hs + (l - 1) * HP * S;
                    dw_block<S, TILE>(dzb + row0 * S, hin + cb * S, 4 * kg, 4 * KG, dWacc[l]);
                    if (NC0 == 2 && l == 0) dw_block<S, TILE>(dzb + row0 * S, hin + (32 + cb) * S, 4 * kg, 4 * KG, dW0x);
                }
                PHASE_MARK(5)  // dW
                if (l > 0) {
                    gemm16<S, HP>(dzb + col0, Wb + (l - 1) * HP * HP + j0, nullptr, HP, dh);
                } else {
                    for (int c0 = 0; c0 < P.Cp; c0 += HP) {
                        if (c0 + j0 < P.Cp) {  // Cp is a multiple of 4: a lane's 4 feature columns are all in or out
                            float2 df[4][2];
                            gemm16_ldw<S>(dzb + col0, W0f + c0 + j0, P.Cp, HP, df);
#pragma unroll
                            for (int o = 0; o < 4; ++o)
                                *reinterpret_cast<float4*>(DF + (c0 + j0 + o) * S + col0) =
                                    make_float4(pick(df, 0, o), pick(df, 1, o), pick(df, 2, o), pick(df, 3, o));
                        }
                    }
                }
            }
        }
        __syncwarp();
        PHASE_MARK(6)  // dh / dfeat GEMMs

        // ---- S: scatter d(features) into the grid gradient, two lanes per sample (warp-local) ----------------------------------
        if (AUX[S + scol] != 0.0f) {
            Corners K;
            make_corners(P, AUX[2 * S + scol], AUX[3 * S + scol], AUX[4 * S + scol], K);
            for (int c4 = 4 * half; c4 < P.Cp; c4 += 8) {
                const float d0 = DF[(c4 + 0) * S + scol], d1 = DF[(c4 + 1) * S + scol];
                const float d2 = DF[(c4 + 2) * S + scol], d3 = DF[(c4 + 3) * S + scol];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float w = K.w[c];
                    if (w != 0.0f) red_add_v4(A.grad_grid + K.off[c] + c4, make_float4(d0 * w, d1 * w, d2 * w, d3 * w));
                }
            }
        }
        PHASE_MARK(7)  // scatter
        __syncthreads();  // the tile's rows are free
        PHASE_MARK(8)  // end-of-tile barrier
    }

    // ---- flush --------------------------------------------------------------------------------------------------------------------
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float v = accWf[o];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (sg == 0) wacc[LMAX * HP + j0 + o] = v;
    }
    {
        float v = accbf, e = loss_part;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, m);
            e += __shfl_xor_sync(0xffffffffu, e, m);
        }
        if (lane == 0) {
            wacc[LMAX * HP + HP] = v;
            wacc[LMAX * HP + HP + 1] = FUSED ? e : 0.0f;
        }
    }
    __syncthreads();
    // fold the per-warp accumulators into warp 0's slot (fixed order)
    for (int e = threadIdx.x; e < kBaccStride; e += blockDim.x) {
        float t = 0.0f;
        for (int w = 0; w < NW; ++w) t += bacc[w * kBaccStride + e];
        bacc[e] = t;
    }
    __syncthreads();
    float* dst = A.partial + ((size_t)blockIdx.x * KGmax + kg) * A.pstride;
    // a row group with fewer warps than KGmax leaves its rows of the last slice to its first warp (zeros)
    float* dstz = (KG < KGmax && kg == 0) ? A.partial + ((size_t)blockIdx.x * KGmax + KGmax - 1) * A.pstride : nullptr;
#pragma unroll
    for (int l = 0; l < LMAX; ++l) {
        if (l < L) {
            const int Kin = l == 0 ? in0 : H;
            const int woff = mlp_w_off(l, in0, H);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int i = row0 + 4 * r;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int col = cb + 8 * c;
                    if (i < H && col < Kin) {
                        dst[woff + i * Kin + col] = dWacc[l][r][c];
                        if (dstz) dstz[woff + i * Kin + col] = 0.0f;
                    }
                    if (NC0 == 2 && l == 0 && i < H && col + 32 < Kin) {
                        dst[woff + i * Kin + col + 32] = dW0x[r][c];
                        if (dstz) dstz[woff + i * Kin + col + 32] = 0.0f;
                    }
                }
            }
        }
    }
    // biases, final layer: slice 0 carries the CTA's sums, the other slices hold zeros there
    for (int k = 0; k < KGmax; ++k) {
        float* d = A.partial + ((size_t)blockIdx.x * KGmax + k) * A.pstride;
        if (threadIdx.x == 0) d[A.pcount] = k == 0 ? bacc[LMAX * HP + HP + 1] : 0.0f;  // loss partial
        for (int l = 0; l < L; ++l)
            for (int j = threadIdx.x; j < H; j += blockDim.x) d[mlp_b_off(l, in0, H) + j] = k == 0 ? bacc[l * HP + j] : 0.0f;
        for (int j = threadIdx.x; j < H + 1; j += blockDim.x)   // Wf gradient (H entries), then the bf gradient (slot HP)
            d[mlp_wf_off(L, in0, H) + j] = k == 0 ? bacc[LMAX * HP + (j < H ? j : HP)] : 0.0f;
    }
    PHASE_MARK(9)  // flush
    PHASE_FLUSH()
}

// Relative cost of a launch: waves x tile / (per-sample efficiency of the width) + a fixed set-up term.  The
// efficiencies are measured (profiles/bwd_sweep.py, B200, C16 H32 L4, n = 262144, relative to 8 warps): 10+ warps are
// register-limited to 168 / 128 registers and spill, fewer than 8 expose the shared-memory latency.
static double tile_cost(int64_t n, int nw, int sms) {
    static const double eff[15] = {0, 0, 0, 0, 0.72, 0.78, 0.86, 0.90, 1.00, 0.85, 0.83, 0.85, 0.90, 0.89, 0.89};
    const int tile = 16 * nw;
    const int64_t tiles = (n + tile - 1) / tile;
    const int64_t waves = (tiles + sms - 1) / sms;
    return (double)waves * tile / eff[nw] + 40.0;
}

static int env_nw() {
    const char* e = getenv("LFGC_BWD_WARPS");  // tuning override
    return e ? atoi(e) : 0;
}

template <int FUSED, int NC0, int NW, int ACT = 0>
static int launch_nw(BwdArgs& A, float* grad_mlp, int accumulate, void* workspace, size_t workspace_bytes,
                     cudaStream_t st) {
    const Layout Lo = make_layout(A.P, NW);
    const size_t smem = (size_t)Lo.total * sizeof(float);
    auto kern = backward_v2_kernel<FUSED, NC0, NW, ACT>;
    LFGC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    constexpr int tile = 16 * NW;
    const int64_t ntiles = (A.n + tile - 1) / tile;
    int64_t grid = sm_count();
    if (grid > ntiles) grid = ntiles;
    constexpr int KGmax = (NW + 3) / 4;
    const size_t need = (size_t)grid * KGmax * A.pstride * sizeof(float);
    if (workspace_bytes < need) return fail(LFGC_E_WORKSPACE, "backward workspace too small: %zu < %zu", workspace_bytes, need);
    A.partial = reinterpret_cast<float*>(workspace);
    (void)launch_pdl(kern, dim3((unsigned)grid), dim3(32 * NW), (size_t)(smem), st, A);
    LFGC_LAUNCH_OK();
    finish_partials(A, (int)grid * KGmax, grad_mlp, accumulate, FUSED ? A.loss_sum : nullptr, st);
    if (!A.defer_reduce) LFGC_LAUNCH_OK();
    return LFGC_OK;
}

static const int kWidths[] = {8, 7, 6, 5, 4};  // ties go to the first; 10-14 warps were measured and never win

template <int FUSED, int NC0>
static int launch(BwdArgs& A, float* grad_mlp, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const int sms = sm_count();
    int best = 0;
    double best_cost = 0.0;
    const int forced = env_nw();
    for (int nw : kWidths) {
        if (forced && nw != forced) continue;
        const Layout Lo = make_layout(A.P, nw);
        if ((size_t)Lo.total * sizeof(float) > (size_t)max_smem_optin()) continue;
        const double c = tile_cost(A.n, nw, sms);
        if (!best || c < best_cost) { best = nw; best_cost = c; }
    }
    switch (best) {
        case 8: return launch_nw<FUSED, NC0, 8>(A, grad_mlp, accumulate, workspace, workspace_bytes, st);
        case 7: return launch_nw<FUSED, NC0, 7>(A, grad_mlp, accumulate, workspace, workspace_bytes, st);
        case 6: return launch_nw<FUSED, NC0, 6>(A, grad_mlp, accumulate, workspace, workspace_bytes, st);
        case 5: return launch_nw<FUSED, NC0, 5>(A, grad_mlp, accumulate, workspace, workspace_bytes, st);
        case 4: return launch_nw<FUSED, NC0, 4>(A, grad_mlp, accumulate, workspace, workspace_bytes, st);
        default: return 1;
    }
}

}  // namespace v2

size_t backward_v2_workspace_floats(int pcount, int sms) { return (size_t)sms * 4 * (size_t)(pcount + 1); }  // <= 16 warps

int launch_backward_v2(BwdArgs& A, int fused, float* grad_mlp, int accumulate, void* workspace, size_t workspace_bytes,
                       cudaStream_t st) {
    const SampleParams& P = A.P;
    if (P.H > v2::HP || P.L > v2::LMAX || P.in0 > 64) return 1;
    if (P.flags & kFlagPlainRelu) {   // Variance_Model: backward-only, ReLU, no grid
        if (fused) return fail(LFGC_E_INVALID, "plain MLP has no fused sampler mode");
        return v2::launch_nw<0, 1, 8, 1>(A, grad_mlp, accumulate, workspace, workspace_bytes, st);
    }
    const bool two = P.in0 > 32;
    if (fused) {
        return two ? v2::launch<1, 2>(A, grad_mlp, accumulate, workspace, workspace_bytes, st)
                   : v2::launch<1, 1>(A, grad_mlp, accumulate, workspace, workspace_bytes, st);
    }
    return two ? v2::launch<0, 2>(A, grad_mlp, accumulate, workspace, workspace_bytes, st)
               : v2::launch<0, 1>(A, grad_mlp, accumulate, workspace, workspace_bytes, st);
}

}  // namespace lfgc

#ifdef LFGC_PHASE_TIMING
extern "C" int lfgc_phase_timing(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (out16) cudaMemcpyFromSymbol(out16, g_phase_cycles, sizeof(unsigned long long) * 16);
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
    }
    return 0;
}
#endif
