// Fused sample kernel (forward): trilinear gather of the channels-last latent grid + Fourier embedding + SnakeAlt
// MLP, one launch.  Replaces model/Feature_Grid_Model.py:62-78 (F.grid_sample, embedder.embed, cat, 4 x
// SnakeAlt(Linear), final Linear, eval clamp) and, in axis-table mode, the tile loop of field_from_net
// (visualization/OutputToVTK.py:11-42).
//
// Work decomposition: a CTA of 4 warps processes tiles of 128 samples.  In the input stage thread t handles sample t
// (8 corners x Cp/4 16-byte loads from the L2-resident grid, sin/cos of the embedding) and writes column t of the
// input block X in shared memory ([feature][sample], row stride 128).  The MLP then runs as a register-tiled
// GEMM per warp: each lane owns 4 samples x HP/4 outputs, reads one 16-byte activation vector and HP/16 16-byte
// weight vectors per k (3 LDS.128 per 32 FFMA at HP=32), applies SnakeAlt in registers and writes the next
// activation block.  All hand-offs are warp-local (__syncwarp), the weights stay in shared memory for the whole
// persistent loop.
#include "sample_common.cuh"

#include <stdlib.h>

namespace lfgc {

// tensor-core variant (sample_forward_tc.cu); returns 1 when the shape is not covered
int launch_forward_tc(const SampleParams& P, const float* coords, const float* const axis[3], int R1, int R2,
                      int64_t first, int64_t n, const float* grid, const float* mlp, float* out, cudaStream_t st);

static bool forward_tc_enabled() {
    const char* e = getenv("LFGC_FORWARD_TC");  // "0" forces the FFMA2 kernel; default: the tcgen05 kernel where it applies
    return e ? (e[0] != '0') : true;
}

struct FwdArgs {
    SampleParams P;
    const float* coords;  // MODE 0: [n][3]
    const float* axis[3]; // MODE 1: per-axis normalised coordinates
    int R1, R2;           // MODE 1: volume extents along dims 1, 2
    int64_t first;        // MODE 1: linear voxel index of the first sample (slab_begin * R1 * R2)
    int64_t n;
    const float* grid;
    const float* mlp;
    float* out;
};

template <int HP>
__host__ __device__ constexpr int fwd_act_rows(int in0p) { return in0p > HP ? in0p : HP; }

template <int HP, int MODE, int ACT>
__global__ void __launch_bounds__(kThreads, HP == 32 ? 6 : 2) sample_forward_kernel(const __grid_constant__ FwdArgs A) {
    LFGC_PDL_PROLOGUE();
    constexpr int S = kTile;
    constexpr int NO = HP / 4;
    extern __shared__ __align__(16) float smem[];
    const SampleParams& P = A.P;
    float* Wt = smem;
    float* bias = Wt + FwdWeights<HP>::total(P.L, P.in0);
    float* Wf = bias + P.L * HP;
    float* bfp = Wf + HP;
    float* act0 = bfp + 4;  // keeps 16-byte alignment: all block sizes above are multiples of 4 floats

    load_fwd_weights<HP>(P, A.mlp, Wt, bias, Wf, bfp);
    __syncthreads();
    const float bf = *bfp;

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const int og = lane >> 3, sg = lane & 7;
    const int col0 = warp * 32 + sg * 4;
    const int j0 = og * NO;
    const bool clampo = (P.flags & LFGC_F_CLAMP) != 0;

    const int64_t ntiles = (A.n + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- input stage: thread = sample --------------------------------------------------------------------
        {
            const int64_t s = tile * kTile + t;
            float cx = 0.f, cy = 0.f, cz = 0.f;
            if (s < A.n) {
                if (MODE == 0) {
                    cx = __ldg(A.coords + 3 * s);
                    cy = __ldg(A.coords + 3 * s + 1);
                    cz = __ldg(A.coords + 3 * s + 2);
                } else {
                    const int64_t v = A.first + s;
                    const int k = (int)(v % A.R2);
                    const int j = (int)((v / A.R2) % A.R1);
                    const int i = (int)(v / ((int64_t)A.R2 * A.R1));
                    cx = __ldg(A.axis[0] + i);
                    cy = __ldg(A.axis[1] + j);
                    cz = __ldg(A.axis[2] + k);
                }
            }
            Corners K;
            make_corners(P, cx, cy, cz, K);
            stage_inputs(P, A.grid, cx, cy, cz, K, act0, S, t);
        }
        __syncwarp();

        // ---- MLP: lane = 4 samples x NO outputs -----------------------------------------------------------------
        // The activations are updated IN PLACE: a lane reads all K rows of its 4 sample columns, the warp syncs, then
        // every lane overwrites its own output rows of those columns.  One buffer instead of a ping-pong pair halves
        // the shared memory per CTA (33 KB at C16/H32), i.e. 6 instead of 4 resident CTAs per SM.
        float y[4] = {0.f, 0.f, 0.f, 0.f};
        for (int l = 0; l < P.L; ++l) {
            float acc[4][NO];
            warp_gemm<HP, S>(act0, Wt + FwdWeights<HP>::layer_off(l, P.in0), bias + l * HP, l == 0 ? P.in0 : HP, col0,
                             j0, acc);
            if (l + 1 < P.L) {
                __syncwarp();  // every lane has finished reading the rows about to be overwritten
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    float4 h;
                    h.x = act_value<ACT>(acc[0][o]);
                    h.y = act_value<ACT>(acc[1][o]);
                    h.z = act_value<ACT>(acc[2][o]);
                    h.w = act_value<ACT>(acc[3][o]);
                    *reinterpret_cast<float4*>(act0 + (j0 + o) * S + col0) = h;
                }
                __syncwarp();
            } else {
                // last hidden layer feeds the final Linear directly from registers (model/Feature_Grid_Model.py:75)
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    const float wf = Wf[j0 + o];
#pragma unroll
                    for (int s = 0; s < 4; ++s) y[s] = fmaf(act_value<ACT>(acc[s][o]), wf, y[s]);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            y[s] += __shfl_xor_sync(0xffffffffu, y[s], 8);
            y[s] += __shfl_xor_sync(0xffffffffu, y[s], 16);
        }
        if (og == 0) {
            const int64_t s0 = tile * kTile + col0;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                float v = y[s] + bf;
                if (clampo) v = fminf(fmaxf(v, -1.0f), 1.0f);
                if (s0 + s < A.n) A.out[s0 + s] = v;
            }
        }
        __syncwarp();  // next tile's input stage overwrites act0
    }
}

template <int HP>
static size_t fwd_smem_bytes(const SampleParams& P) {
    size_t f = (size_t)P.in0 * HP + (size_t)(P.L - 1) * HP * HP + (size_t)P.L * HP + HP + 4;
    f += (size_t)fwd_act_rows<HP>(P.in0p) * kTile;
    return f * sizeof(float);
}

template <int HP, int MODE, int ACT = 0>
static int launch_forward(const FwdArgs& A, cudaStream_t st) {
    const size_t smem = fwd_smem_bytes<HP>(A.P);
    if ((int)smem > max_smem_optin())
        return fail(LFGC_E_UNSUPPORTED, "forward needs %zu B shared memory (> %d)", smem, max_smem_optin());
    auto kern = sample_forward_kernel<HP, MODE, ACT>;
    LFGC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    LFGC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
    if (occ < 1) return fail(LFGC_E_UNSUPPORTED, "forward kernel does not fit on an SM");
    const int64_t ntiles = (A.n + kTile - 1) / kTile;
    int64_t grid = (int64_t)sm_count() * occ;
    if (grid > ntiles) grid = ntiles;
    (void)launch_pdl(kern, dim3((unsigned)grid), dim3(kThreads), (size_t)(smem), st, A);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

}  // namespace lfgc

using namespace lfgc;

extern "C" int64_t lfgc_mlp_param_count(const lfgc_model_desc* m) {
    if (!m) return -1;
    const int64_t in0 = 3 + 6 * m->F + m->C;
    return in0 * m->H + m->H + (int64_t)(m->L - 1) * ((int64_t)m->H * m->H + m->H) + m->H + 1;
}

extern "C" int lfgc_forward(const lfgc_model_desc* m, const float* coords, int64_t n, const float* grid_cl,
                            const float* mlp, float* out, int flags, void* stream) {
    FwdArgs A;
    int rc = fill_sample_params(m, flags, A.P);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!coords || !out)) || !grid_cl || !mlp) return fail(LFGC_E_INVALID, "forward: null pointer or n<0");
    if (n == 0) return LFGC_OK;
    A.coords = coords;
    A.axis[0] = A.axis[1] = A.axis[2] = nullptr;
    A.R1 = A.R2 = 1;
    A.first = 0;
    A.n = n;
    A.grid = grid_cl;
    A.mlp = mlp;
    A.out = out;
    if (forward_tc_enabled()) {
        rc = launch_forward_tc(A.P, coords, nullptr, 1, 1, 0, n, grid_cl, mlp, out, (cudaStream_t)stream);
        if (rc != 1) return rc;
    }
    if (m->H <= 32) return launch_forward<32, 0>(A, (cudaStream_t)stream);
    return launch_forward<64, 0>(A, (cudaStream_t)stream);
}

extern "C" int lfgc_plain_mlp_forward(int H, int L, const float* x, int64_t n, const float* mlp, float* out,
                                      void* stream) {
    FwdArgs A;
    int rc = fill_plain_params(H, L, A.P);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!x || !out)) || !mlp) return fail(LFGC_E_INVALID, "plain_mlp_forward: null pointer or n<0");
    if (n == 0) return LFGC_OK;
    A.coords = x;
    A.axis[0] = A.axis[1] = A.axis[2] = nullptr;
    A.R1 = A.R2 = 1;
    A.first = 0;
    A.n = n;
    A.grid = nullptr;
    A.mlp = mlp;
    A.out = out;
    return launch_forward<32, 0, 1>(A, (cudaStream_t)stream);
}

extern "C" int lfgc_reconstruct(const lfgc_model_desc* m, const float* grid_cl, const float* mlp, const int32_t R[3],
                                const float* axis0, const float* axis1, const float* axis2, int32_t slab_begin,
                                int32_t slab_end, float* out_slab, int flags, void* stream) {
    FwdArgs A;
    int rc = fill_sample_params(m, flags, A.P);
    if (rc) return rc;
    if (!R || !grid_cl || !mlp || !axis0 || !axis1 || !axis2) return fail(LFGC_E_INVALID, "reconstruct: null pointer");
    if (R[0] < 1 || R[1] < 1 || R[2] < 1) return fail(LFGC_E_INVALID, "reconstruct: bad volume extents");
    if (slab_begin < 0 || slab_end > R[0] || slab_begin > slab_end)
        return fail(LFGC_E_INVALID, "reconstruct: slab [%d,%d) outside [0,%d)", slab_begin, slab_end, R[0]);
    if (slab_begin == slab_end) return LFGC_OK;
    if (!out_slab) return fail(LFGC_E_INVALID, "reconstruct: out_slab is null");
    A.coords = nullptr;
    A.axis[0] = axis0; A.axis[1] = axis1; A.axis[2] = axis2;
    A.R1 = R[1]; A.R2 = R[2];
    A.first = (int64_t)slab_begin * R[1] * R[2];
    A.n = (int64_t)(slab_end - slab_begin) * R[1] * R[2];
    A.grid = grid_cl;
    A.mlp = mlp;
    A.out = out_slab;
    if (forward_tc_enabled()) {
        rc = launch_forward_tc(A.P, nullptr, A.axis, A.R1, A.R2, A.first, A.n, grid_cl, mlp, out_slab, (cudaStream_t)stream);
        if (rc != 1) return rc;
    }
    if (m->H <= 32) return launch_forward<32, 1>(A, (cudaStream_t)stream);
    return launch_forward<64, 1>(A, (cudaStream_t)stream);
}
