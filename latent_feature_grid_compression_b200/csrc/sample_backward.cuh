// Argument block shared by the generic (v1) and the wide (v2) fused backward kernels.
#pragma once
#include "sample_common.cuh"

namespace lfgc {

struct BwdArgs {
    SampleParams P;
    // sample source
    const float* coords;        // [n][3] positions: backward-only mode, or host-fed samples in fused mode
    const float* grad_out;      // [n] d(loss)/d(out) in backward-only mode, target values in fused host-fed mode
    const float* volume;        // fused mode
    int R[3];
    float max_idx[3], scales[3];
    unsigned long long n_voxels;
    uint64_t seed, sample_offset, step_stride;
    const int32_t* step_dev;
    const int64_t* explicit_idx;
    float loss_scale2;          // 2 * loss_scale
    float* loss_sum;
    const float* log_sigma;     // fused mode, nullable: per-sample log sigma v_i; d(loss)/d(pred) *= exp(-2 v_i)
    float* dlog_sigma;          // fused mode, nullable: d(loss)/d(v_i) = 2 loss_scale (1 - e_i^2 exp(-2 v_i))
    int64_t n;
    const float* grid;
    const float* mlp;
    float* grad_grid;
    float* partial;             // [gridDim.x][pcount]
    int pcount;                 // packed MLP parameter count
    int pstride;                // floats per workspace slice: pcount + 1 (the last entry carries the loss partial)
    int defer_reduce;           // 1: leave the partial sums in the workspace (lfgc_train_step_partials); the reduction
    int nslices;                //    is folded into lfgc_grid_step, which is told how many slices there are (out)
    float* atomic_out = nullptr;  // lfgc_train_step_accumulate: [pcount + 1] running sums (MLP gradient | loss) the kernel
                                  // ADDS to; the tensor-core kernel does it with atomics from its flush (no workspace
                                  // slices, no reduction launch), the other kernels through the reduction kernel
    const unsigned char* panel_image = nullptr;   // tensor-core kernel: ready-made weight operands (tc_panels.cuh), nullable
    int atomic_slices = 1;        // the tensor-core kernel spreads its CTAs over this many [pcount + 1] slices of atomic_out:
                                  // 148 same-address reductions arriving together serialise in L2 (~3.5 us at one slice)
    // early announcement for lfgc_peer_sum (lfgc_peer_announce): the last CTA to finish stores *ann_epoch + 1 into slot
    // [ann_rank] of every rank's flag array
    int ann_n = 0, ann_rank = 0;
    int* ann_flags[LFGC_MAX_PEERS] = {};
    const int* ann_epoch = nullptr;
    int* ann_ticket = nullptr;
};

// stores the announcement of BwdArgs: called by the first ann_n lanes of ONE converged warp, lane r writes rank r's slot (the
// release stores of the lanes travel in parallel; one thread storing them one after the other cost ~1 us per peer).  What
// the peers go on to read was made visible at device scope -- this GPU's L2, which is where peer reads are served -- before
// the caller decided to announce.
__device__ __forceinline__ void announce_epoch(const BwdArgs& A, int lane) {
    if (lane < A.ann_n) {
        const int e = *reinterpret_cast<const volatile int*>(A.ann_epoch) + 1;
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(A.ann_flags[lane] + A.ann_rank), "r"(e) : "memory");
    }
}
void launch_announce(const BwdArgs& A, cudaStream_t st);

// Tries the wide kernel (v2: 8-12 warps per CTA, S'(z) in registers, packed FFMA2).  Returns LFGC_OK after
// launching, or 1 when the configuration is not covered (the caller then uses the generic kernel), or an error.
int launch_backward_v2(BwdArgs& A, int fused, float* grad_mlp, int accumulate, void* workspace, size_t workspace_bytes,
                       cudaStream_t st);
size_t backward_v2_workspace_floats(int pcount, int sms);

// Tensor-core kernel (sample_backward_tc.cu: tcgen05 3xTF32 for every contraction).  Same return convention.
int launch_backward_tc(BwdArgs& A, int fused, float* grad_mlp, int accumulate, void* workspace, size_t workspace_bytes,
                       cudaStream_t st);

// grad[i] (+)= sum_b partial[b][i] for i < pcount; loss_out[0] (+)= sum_b partial[b][pcount] (fixed order: deterministic);
// accumulate: 0 overwrite, 1 add to grad, 2 add to grad and to loss_out
void launch_reduce_partials(const float* partial, int nslices, int pstride, int pcount, float* grad, int accumulate,
                            float* loss_out, cudaStream_t st);
// what every launcher calls after its kernel: reduce now, or (A.defer_reduce) only record the slice count
inline void finish_partials(BwdArgs& A, int nslices, float* grad, int accumulate, float* loss_out, cudaStream_t st) {
    A.nslices = nslices;
    if (A.atomic_out) {   // accumulate == 2: the loss is added as well
        launch_reduce_partials(A.partial, nslices, A.pstride, A.pcount, A.atomic_out, 2, A.atomic_out + A.pcount, st);
        if (A.ann_n > 0) launch_announce(A, st);
        return;
    }
    if (!A.defer_reduce) launch_reduce_partials(A.partial, nslices, A.pstride, A.pcount, grad, accumulate, loss_out, st);
}

}  // namespace lfgc
