// 1-D building blocks of the separable wavelet synthesis / adjoint (wavelet_transform/Torch_Wavelet_Transform.py:39-57,
// 91-104: grouped conv_transpose3d with stride 2, cropped by `off`), shared by the per-channel grid-step kernel
// (grid_step.cu, pyramid in shared memory) and the large-grid separable kernels (wavelet_sep.cu, passes through HBM).
#pragma once
#include "lfgc_common.cuh"

namespace lfgc {

// n / d for the small non-negative n, d of this kernel (n * d < 2^32) without an integer division: one IMAD.HI
struct FastDiv {
    unsigned mul;   // floor(2^32 / d) + 1; 0 stands for d == 1
};
__host__ __device__ __forceinline__ FastDiv make_fastdiv(unsigned d) {
    FastDiv f;
    f.mul = d <= 1 ? 0u : (unsigned)(0x100000000ull / d) + 1u;
    return f;
}
__host__ __device__ __forceinline__ int fdiv(int n, FastDiv f) {
#ifdef __CUDA_ARCH__
    return f.mul ? (int)__umulhi((unsigned)n, f.mul) : n;
#else
    return f.mul ? (int)(((unsigned long long)(unsigned)n * f.mul) >> 32) : n;
#endif
}

// Work distribution of one pass over `nworkers` threads: ncols columns (1-D lines of the pass) x n positions along the
// filtered dimension; with fewer columns than workers the positions are cut into `groups` ranges of `per`.
// unit u -> group = u / ncols, column = u % ncols (column fastest: neighbouring threads work on neighbouring lines).
struct PassPlan {
    int ncols, n, groups, per;
    FastDiv by_ncols;
};
inline PassPlan make_plan(int ncols, int n, int nworkers) {
    PassPlan p;
    p.ncols = ncols;
    p.n = n;
    int g = ncols > 0 ? nworkers / ncols : 1;
    if (g < 1) g = 1;
    if (g > n) g = n > 0 ? n : 1;
    p.groups = g;
    p.per = (n + g - 1) / g;
    p.by_ncols = make_fastdiv((unsigned)(ncols > 0 ? ncols : 1));
    return p;
}

// CG = true: read through L2 only (ld.global.cg) -- for buffers other CTAs wrote earlier in the same kernel
template <bool CG>
__host__ __device__ __forceinline__ float ld_f(const float* p) {
#ifdef __CUDA_ARCH__
    if (CG) return __ldcg(p);
#endif
    return *p;
}

// synthesis along one axis: inputs LO / HI (band low / high of this axis, element stride si, extent d), outputs OUT
// (stride so, extent t); out[o] = sum_a lo[(o + off) / 2 - a] f_lo[(o + off) % 2 + 2 a] + the same with hi
// (Torch_Wavelet_Transform.py:39-57,91-104: conv_transpose, stride 2, cropped by off)
template <int NT, bool CG = false>
__host__ __device__ __forceinline__ void synth_line(const float* LO, const float* HI, int si, int d, float* OUT, int so,
                                                    int t, int off, int m0, int m1, const float (&flo)[NT],
                                                    const float (&fhi)[NT]) {
    for (int m = m0; m < m1; ++m) {
        float ev = 0.0f, od = 0.0f;
#pragma unroll
        for (int a = 0; a < NT / 2; ++a) {
            const int i = m - a;
            const bool ok = (unsigned)i < (unsigned)d;
            const int ii = ok ? i * si : 0;
            const float l = ok ? ld_f<CG>(LO + ii) : 0.0f;
            const float h = ok ? ld_f<CG>(HI + ii) : 0.0f;
            ev = fmaf(l, flo[2 * a], ev);
            ev = fmaf(h, fhi[2 * a], ev);
            od = fmaf(l, flo[2 * a + 1], od);
            od = fmaf(h, fhi[2 * a + 1], od);
        }
        const int oe = 2 * m - off;
        if ((unsigned)oe < (unsigned)t) OUT[oe * so] = ev;
        if ((unsigned)(oe + 1) < (unsigned)t) OUT[(oe + 1) * so] = od;
    }
}
// adjoint along one axis: g_lo[i] = sum_tt G[2 i + tt - off] f_lo[tt], g_hi likewise, both from the same NT loads
template <int NT, bool CG = false>
__host__ __device__ __forceinline__ void adj_line(const float* G, int sg, int t, int off, float* OLO, float* OHI, int so,
                                                  int i0, int i1, const float (&flo)[NT], const float (&fhi)[NT]) {
    for (int i = i0; i < i1; ++i) {
        float lo = 0.0f, hi = 0.0f;
        const int q0 = 2 * i - off;
#pragma unroll
        for (int tt = 0; tt < NT; ++tt) {
            const int q = q0 + tt;
            const bool ok = (unsigned)q < (unsigned)t;
            const float gq = ok ? ld_f<CG>(G + (ok ? q * sg : 0)) : 0.0f;
            lo = fmaf(gq, flo[tt], lo);
            hi = fmaf(gq, fhi[tt], hi);
        }
        OLO[i * so] = lo;
        OHI[i * so] = hi;
    }
}

}  // namespace lfgc
