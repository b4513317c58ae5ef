// Fused sample kernel (backward / training): recomputes the forward and back-propagates through the MLP, the
// Fourier/concat stage and the trilinear gather in ONE launch; nothing is saved by the forward pass.
// Replaces the autograd backward of model/Feature_Grid_Model.py:62-75 (grid_sampler_3d_backward, cat, addmm,
// sin/pow; training/training.py:137) and, in fused mode, also IndexDataset.__getitem__ (data/IndexDataset.py:90-96),
// the ground-truth lookup (data/Interpolation.py:8-44 at integer positions) and the MSE loss (training.py:130).
//
// Per CTA (4 warps) and tile of 128 samples, everything lives in shared memory as [feature][sample] rows of stride
// 132 floats (conflict-free for the 16-byte row reads of all three GEMM shapes):
//   X (inputs) | h_0..h_{L-2} | g_l = S'(z_l), overwritten in place by dz_l | dfeat
//   G  input stage        thread = sample: (Philox voxel + GT) | coords -> gather + Fourier -> X column
//   F  forward            lane = 4 samples x HP/4 outputs, h and g = S'(z) written per layer
//   B  backward (data)    dz_l = dh_l * g_l ; dh_{l-1} = W_l^T dz_l ; layer 0 only for the feature columns
//   S  scatter            thread = sample: 8 corners x Cp/4 red.global.add.v4.f32 into the grid gradient (L2)
//   W  backward (weights) dW_l = dz_l^T h_{l-1}: warp w owns HP/4 output rows, lane 2x4 block, contraction over
//                         the 128 samples of the tile; accumulated in shared memory across the persistent loop
// At the end each CTA writes its accumulators to a per-CTA slice of the workspace and a second tiny kernel sums the
// slices (deterministic, no atomics on the MLP gradient).
#include "sample_backward.cuh"

#include <stdlib.h>

namespace lfgc {

constexpr int kS = 132;  // row stride (floats) of the activation rows

struct BwdLayout {
    int Wt, bias, Wf, bf, Wb, W0f, acc, X, Hs, Gs, DF, AUX, total;
};

template <int HP>
__host__ __device__ inline BwdLayout bwd_layout(const SampleParams& P, int pcount) {
    BwdLayout o;
    int p = 0;
    o.Wt = p;   p += P.in0 * HP + (P.L - 1) * HP * HP;
    o.bias = p; p += P.L * HP;
    o.Wf = p;   p += HP;
    o.bf = p;   p += 4;
    o.Wb = p;   p += (P.L - 1) * HP * HP;
    o.W0f = p;  p += HP * P.Cp;
    o.acc = p;  p += (pcount + 4) & ~3;    // + 1: loss partial
    o.X = p;    p += P.in0p * kS;
    o.Hs = p;   p += (P.L - 1) * HP * kS;
    o.Gs = p;   p += P.L * HP * kS;
    o.DF = p;   p += P.Cp * kS;
    o.AUX = p;  p += 2 * kS;
    o.total = p;
    return o;
}

// dh[s][o] (o <-> input feature k0+o) = sum_j dz[j][col0+s] * Wb[j][k0+o]; identical shape to the forward GEMM
template <int HP>
__device__ __forceinline__ void warp_gemm_feat(const float* __restrict__ dz, const float* __restrict__ W0f, int Cp,
                                               int col0, int c0, float (&acc)[4][HP / 4]) {
    constexpr int NO = HP / 4;
#pragma unroll
    for (int o = 0; o < NO; ++o)
#pragma unroll
        for (int s = 0; s < 4; ++s) acc[s][o] = 0.0f;
#pragma unroll 4
    for (int j = 0; j < HP; ++j) {
        const float4 a = *reinterpret_cast<const float4*>(dz + j * kS + col0);
#pragma unroll
        for (int o4 = 0; o4 < NO; o4 += 4) {
            if (c0 + o4 < Cp) {  // Cp is a multiple of 4: whole vectors are in or out
                const float4 t = *reinterpret_cast<const float4*>(W0f + j * Cp + c0 + o4);
                const float w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc[0][o4 + q] = fmaf(a.x, w[q], acc[0][o4 + q]);
                    acc[1][o4 + q] = fmaf(a.y, w[q], acc[1][o4 + q]);
                    acc[2][o4 + q] = fmaf(a.z, w[q], acc[2][o4 + q]);
                    acc[3][o4 + q] = fmaf(a.w, w[q], acc[3][o4 + q]);
                }
            }
        }
    }
}

template <int HP, int FUSED>
__global__ void __launch_bounds__(kThreads, 1) sample_backward_kernel(const __grid_constant__ BwdArgs A) {
    LFGC_PDL_PROLOGUE();
    constexpr int NO = HP / 4;        // outputs per lane in the sample-major GEMMs
    constexpr int RPW = HP / 4;       // dW rows per warp
    constexpr int RPL = RPW / 4;      // dW rows per lane
    extern __shared__ __align__(16) float smem[];
    const SampleParams& P = A.P;
    const BwdLayout Lo = bwd_layout<HP>(P, A.pcount);
    float* Wt = smem + Lo.Wt;
    float* bias = smem + Lo.bias;
    float* Wf = smem + Lo.Wf;
    float* Wb = smem + Lo.Wb;
    float* W0f = smem + Lo.W0f;
    float* accum = smem + Lo.acc;
    float* X = smem + Lo.X;
    float* Hs = smem + Lo.Hs;
    float* Gs = smem + Lo.Gs;
    float* DF = smem + Lo.DF;
    float* AUX = smem + Lo.AUX;

    const int H = P.H, in0 = P.in0, L = P.L;
    load_fwd_weights<HP>(P, A.mlp, Wt, bias, Wf, smem + Lo.bf);
    // backward layout: Wb_l[j][k] = W_l[j][k] (l >= 1), W0f[j][c] = W_0[j][3+6F+c]
    load_bwd_weights<HP>(P, A.mlp, Wb, W0f);
    for (int e = threadIdx.x; e < ((A.pcount + 4) & ~3); e += blockDim.x) accum[e] = 0.0f;
    __syncthreads();
    const float bf = smem[Lo.bf];

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const int og = lane >> 3, sg = lane & 7;
    const int col0 = warp * 32 + sg * 4;
    const int j0 = og * NO;
    // dW-phase lane coordinates
    const int wr = lane >> 3;  // 0..3 -> rows warp*RPW + wr + 4*rr
    const int cb = lane & 7;   // columns chunk*32 + cb + 8*c

    float accWf[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) accWf[o] = 0.0f;
    float accbf = 0.0f, loss_part = 0.0f;

    uint64_t sample_base = A.sample_offset;
    if (FUSED && A.step_dev) sample_base += (uint64_t)(*A.step_dev) * A.step_stride;

    const int64_t ntiles = (A.n + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- G: input stage -------------------------------------------------------------------------------------
        Corners K;
        bool valid;
        {
            const int64_t s = tile * kTile + t;
            valid = s < A.n;
            float cx = 0.f, cy = 0.f, cz = 0.f, aux = 0.f;
            if (valid) {
                if (FUSED && A.coords) {  // host-fed step: caller-supplied positions and target values
                    cx = __ldg(A.coords + 3 * s);
                    cy = __ldg(A.coords + 3 * s + 1);
                    cz = __ldg(A.coords + 3 * s + 2);
                    aux = __ldg(A.grad_out + s);
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
                    aux = __ldg(A.volume + v);  // ground truth: exact voxel lookup
                } else {
                    cx = __ldg(A.coords + 3 * s);
                    cy = __ldg(A.coords + 3 * s + 1);
                    cz = __ldg(A.coords + 3 * s + 2);
                    aux = __ldg(A.grad_out + s);
                }
            }
            make_corners(P, cx, cy, cz, K);
            stage_inputs(P, A.grid, cx, cy, cz, K, X, kS, t);
            AUX[t] = aux;
            AUX[kS + t] = valid ? 1.0f : 0.0f;
        }
        __syncwarp();

        // ---- F: forward ---------------------------------------------------------------------------------------------
        float hl[4][NO];  // activations of the last hidden layer stay in registers
        for (int l = 0; l < L; ++l) {
            float acc[4][NO];
            const float* in = l == 0 ? X : Hs + (l - 1) * HP * kS;
            warp_gemm<HP, kS>(in, Wt + FwdWeights<HP>::layer_off(l, in0), bias + l * HP, l == 0 ? in0 : HP, col0, j0, acc);
            float* grow = Gs + l * HP * kS;
            float* hrow = Hs + l * HP * kS;
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                float h[4], g[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) snake_and_grad_precise(acc[s][o], h[s], g[s]);
                *reinterpret_cast<float4*>(grow + (j0 + o) * kS + col0) = make_float4(g[0], g[1], g[2], g[3]);
                if (l + 1 < L) {
                    *reinterpret_cast<float4*>(hrow + (j0 + o) * kS + col0) = make_float4(h[0], h[1], h[2], h[3]);
                } else {
#pragma unroll
                    for (int s = 0; s < 4; ++s) hl[s][o] = h[s];
                }
            }
            __syncwarp();
        }
        float y[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int o = 0; o < NO; ++o) {
            const float wf = Wf[j0 + o];
#pragma unroll
            for (int s = 0; s < 4; ++s) y[s] = fmaf(hl[s][o], wf, y[s]);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            y[s] += __shfl_xor_sync(0xffffffffu, y[s], 8);
            y[s] += __shfl_xor_sync(0xffffffffu, y[s], 16);
        }
        // ---- d loss / d y --------------------------------------------------------------------------------------------
        float dy[4];
        {
            const float4 a4 = *reinterpret_cast<const float4*>(AUX + col0);
            const float4 v4 = *reinterpret_cast<const float4*>(AUX + kS + col0);
            const float aux[4] = {a4.x, a4.y, a4.z, a4.w};
            const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (FUSED) {
                    const float e = (y[s] + bf) - aux[s];
                    dy[s] = vv[s] * A.loss_scale2 * e;
                    if (og == 0) loss_part = fmaf(vv[s] * e, e, loss_part);
                    if (A.log_sigma && vv[s] != 0.0f) {   // Gaussian likelihood with per-sample log sigma
                        const int64_t gs = tile * kTile + col0 + s;
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
        for (int o = 0; o < NO; ++o)
#pragma unroll
            for (int s = 0; s < 4; ++s) accWf[o] = fmaf(dy[s], hl[s][o], accWf[o]);
        if (og == 0) accbf += (dy[0] + dy[1]) + (dy[2] + dy[3]);

        // ---- B: backward through the hidden layers -------------------------------------------------------------------
        float dh[4][NO];
#pragma unroll
        for (int o = 0; o < NO; ++o) {
            const float wf = Wf[j0 + o];
#pragma unroll
            for (int s = 0; s < 4; ++s) dh[s][o] = dy[s] * wf;
        }
        for (int l = L - 1; l >= 0; --l) {
            float* grow = Gs + l * HP * kS;
            float bsum[NO];
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                const float4 g = *reinterpret_cast<const float4*>(grow + (j0 + o) * kS + col0);
                const float4 dz = make_float4(dh[0][o] * g.x, dh[1][o] * g.y, dh[2][o] * g.z, dh[3][o] * g.w);
                *reinterpret_cast<float4*>(grow + (j0 + o) * kS + col0) = dz;
                bsum[o] = (dz.x + dz.y) + (dz.z + dz.w);
            }
            // bias gradient: reduce over the 8 sample-group lanes, one shared-memory atomic per (warp, output)
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                float v = bsum[o];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                if (sg == 0 && j0 + o < H) atomicAdd(accum + mlp_b_off(l, in0, H) + j0 + o, v);
            }
            __syncwarp();
            if (l > 0) {
                warp_gemm<HP, kS>(grow, Wb + (l - 1) * HP * HP, nullptr, HP, col0, j0, dh);
            } else {
                for (int c0 = 0; c0 < P.Cp; c0 += HP) {
                    float df[4][NO];
                    warp_gemm_feat<HP>(grow, W0f, P.Cp, col0, c0 + j0, df);
#pragma unroll
                    for (int o = 0; o < NO; ++o)
                        if (c0 + j0 + o < P.Cp)
                            *reinterpret_cast<float4*>(DF + (c0 + j0 + o) * kS + col0) =
                                make_float4(df[0][o], df[1][o], df[2][o], df[3][o]);
                }
            }
        }
        __syncwarp();

        // ---- S: scatter d(features) into the grid gradient ------------------------------------------------------------
        if (valid) {
            for (int c4 = 0; c4 < P.Cp; c4 += 4) {
                const float d0 = DF[(c4 + 0) * kS + t], d1 = DF[(c4 + 1) * kS + t];
                const float d2 = DF[(c4 + 2) * kS + t], d3 = DF[(c4 + 3) * kS + t];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float w = K.w[c];
                    if (w != 0.0f) red_add_v4(A.grad_grid + K.off[c] + c4, make_float4(d0 * w, d1 * w, d2 * w, d3 * w));
                }
            }
        }
        __syncthreads();

        // ---- W: weight gradients, contraction over the tile's 128 samples ----------------------------------------------
        for (int l = 0; l < L; ++l) {
            const float* dzr = Gs + l * HP * kS;
            const float* hin = l == 0 ? X : Hs + (l - 1) * HP * kS;
            const int Kin = l == 0 ? in0 : H;
            float* wacc = accum + mlp_w_off(l, in0, H);
            for (int chunk = 0; chunk * 32 < Kin; ++chunk) {
                float acc[RPL][4];
#pragma unroll
                for (int r = 0; r < RPL; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
                int colv[4];
                bool colok[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    colv[c] = chunk * 32 + cb + 8 * c;
                    colok[c] = colv[c] < Kin;
                    if (!colok[c]) colv[c] = 0;
                }
#pragma unroll 2
                for (int n0 = 0; n0 < kTile; n0 += 4) {
                    float4 dz[RPL], hv[4];
#pragma unroll
                    for (int r = 0; r < RPL; ++r)
                        dz[r] = *reinterpret_cast<const float4*>(dzr + (warp * RPW + wr + 4 * r) * kS + n0);
#pragma unroll
                    for (int c = 0; c < 4; ++c) hv[c] = *reinterpret_cast<const float4*>(hin + colv[c] * kS + n0);
#pragma unroll
                    for (int r = 0; r < RPL; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            acc[r][c] = fmaf(dz[r].x, hv[c].x, acc[r][c]);
                            acc[r][c] = fmaf(dz[r].y, hv[c].y, acc[r][c]);
                            acc[r][c] = fmaf(dz[r].z, hv[c].z, acc[r][c]);
                            acc[r][c] = fmaf(dz[r].w, hv[c].w, acc[r][c]);
                        }
                }
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    const int i = warp * RPW + wr + 4 * r;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (i < H && colok[c]) wacc[i * Kin + colv[c]] += acc[r][c];
                }
            }
        }
        __syncthreads();
    }

    // ---- flush: per-CTA partial sums in the global packed layout --------------------------------------------------------
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        float v = accWf[o];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (sg == 0 && j0 + o < H) atomicAdd(accum + mlp_wf_off(L, in0, H) + j0 + o, v);
    }
    {
        float v = accbf, e = loss_part;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, m);
            e += __shfl_xor_sync(0xffffffffu, e, m);
        }
        if (lane == 0) {
            atomicAdd(accum + mlp_wf_off(L, in0, H) + H, v);
            if (FUSED) atomicAdd(accum + A.pcount, e);
        }
    }
    __syncthreads();
    float* dst = A.partial + (size_t)blockIdx.x * A.pstride;
    for (int e = threadIdx.x; e < A.pstride; e += blockDim.x) dst[e] = accum[e];
}

// grad[i] (+)= sum over CTA slices; block = 32 entries x 32 slice lanes (fixed summation order: deterministic).
// Entry pcount of every slice is the loss partial; its sum goes to loss_out (overwritten).
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ partial, int nslices,
                                                               int pstride, int pcount, float* __restrict__ grad,
                                                               int accumulate, float* __restrict__ loss_out) {
    LFGC_PDL_PROLOGUE();
    __shared__ float red[32][33];
    const int px = threadIdx.x & 31, sy = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + px;
    float s = 0.0f;
    if (i <= pcount) {
        // independent loads first (one L2 round trip for up to 4 x 32 slices), then the ordered sum
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int b = sy + 32 * u;
            v[u] = b < nslices ? partial[(size_t)b * pstride + i] : 0.0f;
        }
        s = (v[0] + v[1]) + (v[2] + v[3]);
        for (int b = sy + 128; b < nslices; b += 32) s += partial[(size_t)b * pstride + i];
    }
    red[sy][px] = s;
    __syncthreads();
    if (sy == 0 && i <= pcount) {
        float t = 0.0f;
#pragma unroll
        for (int r = 0; r < 32; ++r) t += red[r][px];
        if (i < pcount) grad[i] = accumulate ? grad[i] + t : t;
        else if (loss_out) loss_out[0] = accumulate == 2 ? loss_out[0] + t : t;
    }
}

__global__ void announce_kernel(const __grid_constant__ BwdArgs A) {
    LFGC_PDL_PROLOGUE();
    announce_epoch(A, (int)threadIdx.x);
}
void launch_announce(const BwdArgs& A, cudaStream_t st) {
    (void)launch_pdl(announce_kernel, dim3(1), dim3(32), (size_t)0, st, A);
}

void launch_reduce_partials(const float* partial, int nslices, int pstride, int pcount, float* grad, int accumulate,
                            float* loss_out, cudaStream_t st) {
    (void)launch_pdl(reduce_partials_kernel, dim3((pcount + 1 + 31) / 32), dim3(1024), (size_t)(0), st, partial, nslices, pstride, pcount, grad, accumulate,
                                                                  loss_out);
}

template <int HP, int FUSED>
static int launch_backward(BwdArgs& A, float* grad_mlp, int accumulate, void* workspace, size_t workspace_bytes,
                           cudaStream_t st) {
    {
        // tensor-core kernel first (LFGC_BACKWARD_TC=0 forces the FFMA2 kernels); it covers the MSE, log-likelihood and
        // backward-only modes of the SnakeAlt model (the ReLU Variance_Model stays on the FFMA2 kernel)
        const char* e = getenv("LFGC_BACKWARD_TC");
        if (!(e && e[0] == '0') && !(A.P.flags & kFlagPlainRelu)) {
            const int rc = launch_backward_tc(A, FUSED, grad_mlp, accumulate, workspace, workspace_bytes, st);
            if (rc != 1) return rc;
        }
    }
    {
        const int rc = launch_backward_v2(A, FUSED, grad_mlp, accumulate, workspace, workspace_bytes, st);
        if (rc != 1) return rc;
    }
    if (A.P.flags & kFlagPlainRelu) return fail(LFGC_E_UNSUPPORTED, "plain MLP shape not covered by the wide kernel");
    const BwdLayout Lo = bwd_layout<HP>(A.P, A.pcount);
    const size_t smem = (size_t)Lo.total * sizeof(float);
    if ((int)smem > max_smem_optin())
        return fail(LFGC_E_UNSUPPORTED, "backward needs %zu B shared memory (> %d): C=%d H=%d L=%d F=%d", smem,
                    max_smem_optin(), A.P.C, A.P.H, A.P.L, A.P.F);
    auto kern = sample_backward_kernel<HP, FUSED>;
    LFGC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (A.n + kTile - 1) / kTile;
    int64_t grid = sm_count();
    if (grid > ntiles) grid = ntiles;
    if (workspace_bytes < (size_t)grid * A.pstride * sizeof(float))
        return fail(LFGC_E_WORKSPACE, "backward workspace too small: %zu < %zu", workspace_bytes,
                    (size_t)grid * A.pstride * sizeof(float));
    A.partial = reinterpret_cast<float*>(workspace);
    (void)launch_pdl(kern, dim3((unsigned)grid), dim3(kThreads), (size_t)(smem), st, A);
    LFGC_LAUNCH_OK();
    finish_partials(A, (int)grid, grad_mlp, accumulate, FUSED ? A.loss_sum : nullptr, st);
    if (!A.defer_reduce) LFGC_LAUNCH_OK();
    return LFGC_OK;
}

}  // namespace lfgc

using namespace lfgc;

extern "C" size_t lfgc_backward_workspace_bytes(const lfgc_model_desc* m) {
    if (!m) return 0;
    const int64_t p = lfgc_mlp_param_count(m);
    int sms = sm_count();
    if (sms <= 0) sms = 148;  // no device visible (build container): size for a B200
    size_t fl = (size_t)sms * (size_t)(p + 1);
    const size_t v2 = backward_v2_workspace_floats((int)p, sms);
    if (v2 > fl) fl = v2;
    return fl * sizeof(float);
}

static int common_checks(const lfgc_model_desc* m, int64_t n, const float* grid_cl, const float* mlp,
                         float* grad_grid_cl, float* grad_mlp, void* workspace) {
    if (n < 0 || !grid_cl || !mlp || !grad_grid_cl || !grad_mlp || !workspace)
        return fail(LFGC_E_INVALID, "backward: null pointer or n<0");
    if (m->H > 32) return fail(LFGC_E_UNSUPPORTED, "backward supports n_hidden_size <= 32 (got %d)", m->H);
    return LFGC_OK;
}

extern "C" int lfgc_backward(const lfgc_model_desc* m, const float* coords, int64_t n, const float* grad_out,
                             const float* grid_cl, const float* mlp, float* grad_grid_cl, float* grad_mlp,
                             float* grad_coords, int accumulate_mlp, void* workspace, size_t workspace_bytes,
                             void* stream) {
    BwdArgs A;
    A.defer_reduce = 0;
    A.nslices = 0;
    int rc = fill_sample_params(m, 0, A.P);
    if (rc) return rc;
    rc = common_checks(m, n, grid_cl, mlp, grad_grid_cl, grad_mlp, workspace);
    if (rc) return rc;
    if (grad_coords) return fail(LFGC_E_UNSUPPORTED, "coordinate gradients are not produced (the reference never reads them)");
    if (n > 0 && (!coords || !grad_out)) return fail(LFGC_E_INVALID, "backward: coords/grad_out null");
    A.pcount = (int)lfgc_mlp_param_count(m);
    A.pstride = A.pcount + 1;
    if (n == 0) {
        if (!accumulate_mlp) LFGC_CUDA_OK(cudaMemsetAsync(grad_mlp, 0, A.pcount * sizeof(float), (cudaStream_t)stream));
        return LFGC_OK;
    }
    A.coords = coords;
    A.grad_out = grad_out;
    A.volume = nullptr;
    A.explicit_idx = nullptr;
    A.loss_sum = nullptr;
    A.log_sigma = nullptr;
    A.dlog_sigma = nullptr;
    A.loss_scale2 = 0.0f;
    A.seed = A.sample_offset = A.step_stride = 0;
    A.step_dev = nullptr;
    A.n_voxels = 1;
    for (int a = 0; a < 3; ++a) { A.R[a] = 1; A.max_idx[a] = 1.0f; A.scales[a] = 1.0f; }
    A.n = n;
    A.grid = grid_cl;
    A.mlp = mlp;
    A.grad_grid = grad_grid_cl;
    return launch_backward<32, 0>(A, grad_mlp, accumulate_mlp, workspace, workspace_bytes, (cudaStream_t)stream);
}

static int train_step_impl(const lfgc_model_desc* m, const float* volume, const int32_t R[3], int64_t n,
                           uint64_t seed, uint64_t sample_offset, const int32_t* step_dev, uint64_t step_stride,
                           const int64_t* explicit_idx, const float* explicit_coords, const float* explicit_gt,
                           float loss_scale, const float* log_sigma, float* dlog_sigma, const float* grid_cl,
                           const float* mlp, float* grad_grid_cl, float* grad_mlp, float* loss_sum, int accumulate_mlp,
                           void* workspace, size_t workspace_bytes, void* stream, int32_t* nslices_out = nullptr,
                           float* atomic_out = nullptr, const lfgc_peer_announce* announce = nullptr,
                           int atomic_slices = 1, const float* tc_panels = nullptr) {
    BwdArgs A;
    A.defer_reduce = nslices_out ? 1 : 0;
    A.nslices = 0;
    A.atomic_out = atomic_out;
    A.atomic_slices = atomic_slices;
    if (tc_panels && (reinterpret_cast<uintptr_t>(tc_panels) & 15))
        return fail(LFGC_E_INVALID, "train_step: the panel image must be 16-byte aligned");
    A.panel_image = reinterpret_cast<const unsigned char*>(tc_panels);
    if (announce) {
        if (!atomic_out || announce->n_peers < 1 || announce->n_peers > LFGC_MAX_PEERS || announce->rank < 0 ||
            announce->rank >= announce->n_peers || !announce->epoch || !announce->ticket)
            return fail(LFGC_E_INVALID, "train_step_accumulate: bad announcement block");
        A.ann_n = announce->n_peers;
        A.ann_rank = announce->rank;
        A.ann_epoch = announce->epoch;
        A.ann_ticket = announce->ticket;
        for (int r = 0; r < announce->n_peers; ++r) {
            if (!announce->flags[r]) return fail(LFGC_E_INVALID, "train_step_accumulate: flag array of rank %d is null", r);
            A.ann_flags[r] = announce->flags[r];
        }
    }
    int rc = fill_sample_params(m, 0, A.P);
    if (rc) return rc;
    if (nslices_out) {
        *nslices_out = 0;
        grad_mlp = reinterpret_cast<float*>(workspace);   // unused in deferred mode; keeps the common checks uniform
    }
    if (atomic_out) {
        grad_mlp = atomic_out;
        accumulate_mlp = 1;
    }
    rc = common_checks(m, n, grid_cl, mlp, grad_grid_cl, grad_mlp, workspace);
    if (rc) return rc;
    if ((explicit_coords == nullptr) != (explicit_gt == nullptr))
        return fail(LFGC_E_INVALID, "train_step: explicit_coords and explicit_gt go together");
    if (!explicit_coords && (!volume || !R || R[0] < 1 || R[1] < 1 || R[2] < 1))
        return fail(LFGC_E_INVALID, "train_step: bad volume");
    A.pcount = (int)lfgc_mlp_param_count(m);
    A.pstride = A.pcount + 1;
    if (n == 0) {
        if (!accumulate_mlp) LFGC_CUDA_OK(cudaMemsetAsync(grad_mlp, 0, A.pcount * sizeof(float), (cudaStream_t)stream));
        return LFGC_OK;
    }
    A.coords = explicit_coords;   // non-null: host-fed samples (positions) ...
    A.grad_out = explicit_gt;     // ... and their target values
    A.volume = volume;
    A.explicit_idx = explicit_idx;
    A.loss_sum = loss_sum;
    A.log_sigma = log_sigma;
    A.dlog_sigma = dlog_sigma;
    A.loss_scale2 = 2.0f * loss_scale;
    A.seed = seed;
    A.sample_offset = sample_offset;
    A.step_dev = step_dev;
    A.step_stride = step_stride;
    const int32_t unit[3] = {1, 1, 1};
    if (!R) R = unit;
    A.n_voxels = (unsigned long long)R[0] * R[1] * R[2];
    float mx = 0.0f;
    for (int a = 0; a < 3; ++a) {
        A.R[a] = R[a];
        A.max_idx[a] = (float)R[a] - 1.0f;  // vol_res - 1 in fp32 (data/IndexDataset.py:57)
        if (A.max_idx[a] > mx) mx = A.max_idx[a];
    }
    for (int a = 0; a < 3; ++a) A.scales[a] = mx > 0.0f ? A.max_idx[a] / mx : 1.0f;  // fp32 division (:65)
    A.n = n;
    A.grid = grid_cl;
    A.mlp = mlp;
    A.grad_grid = grad_grid_cl;
    rc = launch_backward<32, 1>(A, grad_mlp, accumulate_mlp, workspace, workspace_bytes, (cudaStream_t)stream);
    if (rc == LFGC_OK && nslices_out) *nslices_out = A.nslices;
    return rc;
}

extern "C" int lfgc_train_step(const lfgc_model_desc* m, const float* volume, const int32_t R[3], int64_t n,
                               uint64_t seed, uint64_t sample_offset, const int32_t* step_dev, uint64_t step_stride,
                               const int64_t* explicit_idx, const float* explicit_coords, const float* explicit_gt,
                               float loss_scale, const float* grid_cl, const float* mlp, float* grad_grid_cl, float* grad_mlp,
                               float* loss_sum, int accumulate_mlp, void* workspace, size_t workspace_bytes,
                               void* stream) {
    return train_step_impl(m, volume, R, n, seed, sample_offset, step_dev, step_stride, explicit_idx, explicit_coords,
                           explicit_gt, loss_scale, nullptr, nullptr, grid_cl, mlp, grad_grid_cl, grad_mlp, loss_sum,
                           accumulate_mlp, workspace, workspace_bytes, stream);
}

extern "C" int lfgc_train_step_partials(const lfgc_model_desc* m, const float* volume, const int32_t R[3], int64_t n,
                                        uint64_t seed, uint64_t sample_offset, const int32_t* step_dev,
                                        uint64_t step_stride, const int64_t* explicit_idx, const float* explicit_coords,
                                        const float* explicit_gt, float loss_scale, const float* grid_cl, const float* mlp,
                                        float* grad_grid_cl, const float* tc_panels, void* workspace,
                                        size_t workspace_bytes, int32_t* nslices_out, void* stream) {
    if (!nslices_out) return fail(LFGC_E_INVALID, "train_step_partials: nslices_out is null");
    if (n == 0) return fail(LFGC_E_INVALID, "train_step_partials: n must be positive (nothing would be written)");
    return train_step_impl(m, volume, R, n, seed, sample_offset, step_dev, step_stride, explicit_idx, explicit_coords,
                           explicit_gt, loss_scale, nullptr, nullptr, grid_cl, mlp, grad_grid_cl, nullptr, nullptr, 0,
                           workspace, workspace_bytes, stream, nslices_out, nullptr, nullptr, 1, tc_panels);
}

extern "C" int lfgc_train_step_accumulate(const lfgc_model_desc* m, const float* volume, const int32_t R[3], int64_t n,
                                          uint64_t seed, uint64_t sample_offset, const int32_t* step_dev,
                                          uint64_t step_stride, const int64_t* explicit_idx, const float* explicit_coords,
                                          const float* explicit_gt, float loss_scale, const float* grid_cl,
                                          const float* mlp, float* grad_grid_cl, float* grad_mlp_loss, int n_slices,
                                          const lfgc_peer_announce* announce, const float* tc_panels, void* workspace,
                                          size_t workspace_bytes, void* stream) {
    if (!grad_mlp_loss) return fail(LFGC_E_INVALID, "train_step_accumulate: grad_mlp_loss is null");
    if (n_slices < 1 || n_slices > 1024) return fail(LFGC_E_INVALID, "train_step_accumulate: n_slices = %d", n_slices);
    if (n == 0 && announce) return fail(LFGC_E_INVALID, "train_step_accumulate: n must be positive with an announcement");
    return train_step_impl(m, volume, R, n, seed, sample_offset, step_dev, step_stride, explicit_idx, explicit_coords,
                           explicit_gt, loss_scale, nullptr, nullptr, grid_cl, mlp, grad_grid_cl, nullptr, nullptr, 1,
                           workspace, workspace_bytes, stream, nullptr, grad_mlp_loss, announce, n_slices, tc_panels);
}

extern "C" int lfgc_train_step_weighted(const lfgc_model_desc* m, const float* volume, const int32_t R[3], int64_t n,
                                        uint64_t seed, uint64_t sample_offset, const int32_t* step_dev,
                                        uint64_t step_stride, const int64_t* explicit_idx, const float* explicit_coords,
                                        const float* explicit_gt, float loss_scale, const float* log_sigma,
                                        float* dlog_sigma, const float* grid_cl, const float* mlp, float* grad_grid_cl,
                                        float* grad_mlp, float* loss_sum, int accumulate_mlp, void* workspace,
                                        size_t workspace_bytes, void* stream) {
    return train_step_impl(m, volume, R, n, seed, sample_offset, step_dev, step_stride, explicit_idx, explicit_coords,
                           explicit_gt, loss_scale, log_sigma, dlog_sigma, grid_cl, mlp, grad_grid_cl, grad_mlp, loss_sum,
                           accumulate_mlp, workspace, workspace_bytes, stream);
}

static int64_t plain_mlp_params(int H, int L) { return 3 * (int64_t)H + H + (int64_t)(L - 1) * ((int64_t)H * H + H) + H + 1; }

extern "C" size_t lfgc_plain_mlp_workspace_bytes(int H, int L) {
    if (H < 1 || L < 1) return 0;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    return backward_v2_workspace_floats((int)plain_mlp_params(H, L), sms) * sizeof(float);
}

extern "C" int lfgc_plain_mlp_backward(int H, int L, const float* x, int64_t n, const float* grad_out, const float* mlp,
                                       float* grad_mlp, int accumulate, void* workspace, size_t workspace_bytes,
                                       void* stream) {
    BwdArgs A;
    A.defer_reduce = 0;
    A.nslices = 0;
    int rc = fill_plain_params(H, L, A.P);
    if (rc) return rc;
    if (n < 0 || !mlp || !grad_mlp || !workspace) return fail(LFGC_E_INVALID, "plain_mlp_backward: null pointer or n<0");
    if (n > 0 && (!x || !grad_out)) return fail(LFGC_E_INVALID, "plain_mlp_backward: x/grad_out null");
    A.pcount = (int)plain_mlp_params(H, L);
    A.pstride = A.pcount + 1;
    if (n == 0) {
        if (!accumulate) LFGC_CUDA_OK(cudaMemsetAsync(grad_mlp, 0, A.pcount * sizeof(float), (cudaStream_t)stream));
        return LFGC_OK;
    }
    A.coords = x;
    A.grad_out = grad_out;
    A.volume = nullptr;
    A.explicit_idx = nullptr;
    A.loss_sum = nullptr;
    A.log_sigma = nullptr;
    A.dlog_sigma = nullptr;
    A.loss_scale2 = 0.0f;
    A.seed = A.sample_offset = A.step_stride = 0;
    A.step_dev = nullptr;
    A.n_voxels = 1;
    for (int a = 0; a < 3; ++a) { A.R[a] = 1; A.max_idx[a] = 1.0f; A.scales[a] = 1.0f; }
    A.n = n;
    A.grid = nullptr;
    A.mlp = mlp;
    A.grad_grid = nullptr;
    return launch_backward<32, 0>(A, grad_mlp, accumulate, workspace, workspace_bytes, (cudaStream_t)stream);
}
