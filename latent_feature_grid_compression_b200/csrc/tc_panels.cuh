// Weight operands of the tensor-core training kernel (sample_backward_tc.cu) as a ready-made IMAGE in global memory.
//
// Every launch of the kernel used to rebuild, per CTA, the tf32 hi / lo split of all weights in the two shared-memory panel
// layouts (forward: B[n = j][K = k], backward: B[n = k][K = j]) from the packed parameter block: zero fill, an L2 round
// trip for the parameters and ~13 conflicted shared-memory stores per thread -- 12 k of the 64 k cycles of a launch at
// 32768 samples.  The weights only change in the optimiser step, so the thread that applies Adam to a parameter
// (lfgc_grid_step's MLP CTAs) also stores its hi / lo parts at their panel positions here, and the training kernel's setup
// becomes two straight 128-bit copies.  Layout (bytes):
//     [0, 512)            bias[l * 32 + j]            (LMAX = 4 layers x HP = 32, zero padded)
//     [512, 656)          final layer: wf[0..31], bf at [32]
//     [1024, ...)         WfHi | WfLo | WbHi | WbLo   exactly as in the kernel's shared memory
// Pad entries are never written: the image is zero-initialised once by its owner.
#pragma once
#include "sample_common.cuh"

namespace lfgc {

struct TcPanelMap {
    int in0, H, L, Cp, nfix, K0p, Np0;
    int wfBytes, wbBytes, wb0;       // bytes of one of hi / lo; bytes of the layer-0 backward operand
    int total_bytes;                 // 0: the tensor-core kernel does not cover this model
};

constexpr int kTcHP = 32, kTcLMAX = 4, kTcPanelW = kTcHP * 16;
constexpr int kTcHdrBias = 0, kTcHdrWf = kTcLMAX * kTcHP * 4, kTcHdrCopy = kTcHdrWf + (kTcHP + 4) * 4, kTcHdrBytes = 1024;

__host__ __device__ inline TcPanelMap make_tc_panel_map(const SampleParams& P) {
    TcPanelMap M;
    M.in0 = P.in0; M.H = P.H; M.L = P.L; M.Cp = P.Cp;
    M.nfix = 3 + 6 * P.F;
    const int K0 = P.Cp + M.nfix;
    M.K0p = (K0 + 7) & ~7;
    M.Np0 = (P.Cp + 15) & ~15;
    M.wb0 = (kTcHP / 4) * M.Np0 * 16;
    M.wfBytes = (M.K0p + (P.L - 1) * kTcHP) / 4 * kTcPanelW;
    M.wbBytes = M.wb0 + (P.L - 1) * (kTcHP / 4) * kTcPanelW;
    const bool ok = P.H <= kTcHP && P.L >= 1 && P.L <= kTcLMAX && M.K0p <= 56 && !(P.flags & kFlagPlainRelu);
    M.total_bytes = ok ? kTcHdrBytes + 2 * M.wfBytes + 2 * M.wbBytes : 0;
    return M;
}

__host__ __device__ __forceinline__ void tc_split_tf32(float x, float& hi, float& lo) {
#ifdef __CUDA_ARCH__
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);  // round to nearest tf32
#else
    union { float f; unsigned u; } v; v.f = x; v.u = (v.u + 0x1000u) & 0xFFFFE000u; hi = v.f;
#endif
    lo = x - hi;
}

// Store parameter `idx` of the packed block (value `x`) at its positions in the image.
__device__ __forceinline__ void tc_panel_store(unsigned char* img, const TcPanelMap& M, int idx, float x) {
    const int H = M.H, in0 = M.in0;
    const int wfo = mlp_wf_off(M.L, in0, H);
    if (idx >= wfo) {   // final layer
        if (idx <= wfo + H) reinterpret_cast<float*>(img + kTcHdrWf)[idx == wfo + H ? kTcHP : idx - wfo] = x;
        return;
    }
    int l = 0;
    while (l + 1 < M.L && idx >= mlp_w_off(l + 1, in0, H)) ++l;
    const int K = l == 0 ? in0 : H;
    const int e = idx - mlp_w_off(l, in0, H);
    if (e >= K * H) {   // bias
        reinterpret_cast<float*>(img + kTcHdrBias)[l * kTcHP + (e - K * H)] = x;
        return;
    }
    const int j = e / K, r = e - j * K;
    int k = r;
    if (l == 0) k = r < M.nfix ? M.Cp + r : r - M.nfix;   // permuted layer-0 columns: [features | xyz | Fourier]
    float hi, lo;
    tc_split_tf32(x, hi, lo);
    unsigned char* panels = img + kTcHdrBytes;
    const int fbase = (l == 0 ? 0 : (M.K0p + (l - 1) * kTcHP) / 4) * kTcPanelW;
    const int fo = fbase + (k >> 2) * kTcPanelW + j * 16 + (k & 3) * 4;                 // forward: B[n = j][K = k]
    *reinterpret_cast<float*>(panels + fo) = hi;
    *reinterpret_cast<float*>(panels + M.wfBytes + fo) = lo;
    if (l > 0 || k < M.Cp) {                                                            // backward: B[n = k][K = j]
        const int bbase = l == 0 ? 0 : M.wb0 + (l - 1) * (kTcHP / 4) * kTcPanelW;
        const int brows = l == 0 ? M.Np0 : kTcHP;
        const int bo = bbase + (j >> 2) * brows * 16 + k * 16 + (j & 3) * 4;
        *reinterpret_cast<float*>(panels + 2 * M.wfBytes + bo) = hi;
        *reinterpret_cast<float*>(panels + 2 * M.wfBytes + M.wbBytes + bo) = lo;
    }
}

}  // namespace lfgc
