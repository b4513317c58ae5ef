// Voxel sampler, ground-truth lookup and deviation statistics.
//   lfgc_sample     replaces IndexDataset.__getitem__ + DataLoader collation (data/IndexDataset.py:90-96) and the
//                   ground-truth lookup at integer positions (training/training.py:107-109)
//   lfgc_trilinear  replaces trilinear_f_interpolation (data/Interpolation.py:8-44) for arbitrary positions
//   lfgc_deviation_stats  replaces calculate_deviation_statistics (visualization/OutputToVTK.py:53-60)
#include "lfgc_common.cuh"

namespace lfgc {

struct SampleArgs {
    const float* volume;
    int R[3];
    float max_idx[3], scales[3];
    unsigned long long n_voxels;
    int64_t n;
    uint64_t seed, offset, step_stride;
    const int32_t* step_dev;
    const int64_t* explicit_idx;
    float* raw;
    float* norm;
    float* gt;
};

__global__ void sample_kernel(const __grid_constant__ SampleArgs A) {
    LFGC_PDL_PROLOGUE();
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.n) return;
    uint64_t base = A.offset;
    if (A.step_dev) base += (uint64_t)(*A.step_dev) * A.step_stride;
    const unsigned long long v = A.explicit_idx ? (unsigned long long)A.explicit_idx[s]
                                                : philox_voxel(A.seed, base + (uint64_t)s, A.n_voxels);
    const unsigned long long r12 = (unsigned long long)A.R[1] * A.R[2];
    const int i = (int)(v / r12);
    const int j = (int)((v / A.R[2]) % A.R[1]);
    const int k = (int)(v % A.R[2]);
    if (A.raw) {
        A.raw[3 * s] = (float)i;
        A.raw[3 * s + 1] = (float)j;
        A.raw[3 * s + 2] = (float)k;
    }
    if (A.norm) {
        A.norm[3 * s] = normalized_coord((float)i, A.max_idx[0], A.scales[0]);
        A.norm[3 * s + 1] = normalized_coord((float)j, A.max_idx[1], A.scales[1]);
        A.norm[3 * s + 2] = normalized_coord((float)k, A.max_idx[2], A.scales[2]);
    }
    if (A.gt) A.gt[s] = __ldg(A.volume + v);
}

struct TriArgs {
    const float* p;
    const float* f;
    int R[3];
    float min_bb[3], max_bb[3];
    int64_t n;
    float* out;
};

__global__ void trilinear_kernel(const __grid_constant__ TriArgs A) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.n) return;
    int lo[3], hi[3];
    float al[3], oal[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float res1 = __fsub_rn((float)A.R[a], 1.0f);
        // ((p - min_bb) / (max_bb - min_bb)) * (res - 1), fp32 (Interpolation.py:10)
        const float np_ = __fmul_rn(__fdiv_rn(__fsub_rn(A.p[3 * s + a], A.min_bb[a]), __fsub_rn(A.max_bb[a], A.min_bb[a])), res1);
        const float fl = floorf(np_), ce = ceilf(np_);
        const double diff = fmax((double)ce - (double)fl, 1e-12);              // :15-16
        const float alpha = (float)(((double)np_ - (double)fl) / diff);        // :19-20
        al[a] = alpha;
        oal[a] = __fsub_rn(1.0f, alpha);
        // the reference raises on out-of-range positions; clamp instead of faulting
        lo[a] = min(max((int)fl, 0), A.R[a] - 1);
        hi[a] = min(max((int)ce, 0), A.R[a] - 1);
    }
    auto at = [&](int x, int y, int z) { return __ldg(A.f + ((int64_t)x * A.R[1] + y) * A.R[2] + z); };
    auto lerp = [](float oa, float a, float v0, float v1) { return __fadd_rn(__fmul_rn(oa, v0), __fmul_rn(a, v1)); };
    const float x00 = lerp(oal[0], al[0], at(lo[0], lo[1], lo[2]), at(hi[0], lo[1], lo[2]));
    const float x10 = lerp(oal[0], al[0], at(lo[0], hi[1], lo[2]), at(hi[0], hi[1], lo[2]));
    const float x01 = lerp(oal[0], al[0], at(lo[0], lo[1], hi[2]), at(hi[0], lo[1], hi[2]));
    const float x11 = lerp(oal[0], al[0], at(lo[0], hi[1], hi[2]), at(hi[0], hi[1], hi[2]));
    const float y0 = lerp(oal[1], al[1], x00, x10);
    const float y1 = lerp(oal[1], al[1], x01, x11);
    A.out[s] = lerp(oal[2], al[2], y0, y1);
}

__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) >= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) <= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}

__global__ void deviation_stats_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int64_t n,
                                       double* acc) {
    double sq = 0.0, ab = 0.0, mx = -INFINITY, mn = INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = gt[i];
        const float d = g - pred[i];  // fp32 difference as torch computes it; accumulation in fp64
        sq += (double)d * (double)d;
        ab += fabs((double)d);
        mx = fmax(mx, (double)g);
        mn = fmin(mn, (double)g);
    }
    for (int m = 16; m >= 1; m >>= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, m);
        ab += __shfl_xor_sync(0xffffffffu, ab, m);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, m));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(acc + 0, sq);
        atomicAdd(acc + 1, ab);
        atomic_max_double(acc + 2, mx);
        atomic_min_double(acc + 3, mn);
    }
}

}  // namespace lfgc

using namespace lfgc;

extern "C" int lfgc_sample_stream(const float* volume, const int32_t R[3], int64_t n, uint64_t seed,
                                  uint64_t sample_offset, const int32_t* step_dev, uint64_t step_stride,
                                  const int64_t* explicit_idx, float* raw_out, float* norm_out, float* gt_out,
                                  void* stream) {
    if (!R || R[0] < 1 || R[1] < 1 || R[2] < 1 || n < 0) return fail(LFGC_E_INVALID, "sample: bad arguments");
    if (gt_out && !volume) return fail(LFGC_E_INVALID, "sample: gt requested without a volume");
    if (n == 0) return LFGC_OK;
    SampleArgs A;
    A.volume = volume;
    float mx = 0.0f;
    for (int a = 0; a < 3; ++a) {
        A.R[a] = R[a];
        A.max_idx[a] = (float)R[a] - 1.0f;
        if (A.max_idx[a] > mx) mx = A.max_idx[a];
    }
    for (int a = 0; a < 3; ++a) A.scales[a] = A.max_idx[a] / mx;
    A.n_voxels = (unsigned long long)R[0] * R[1] * R[2];
    A.n = n;
    A.seed = seed;
    A.offset = sample_offset;
    A.step_dev = step_dev;
    A.step_stride = step_stride;
    A.explicit_idx = explicit_idx;
    A.raw = raw_out;
    A.norm = norm_out;
    A.gt = gt_out;
    (void)launch_pdl(sample_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, A);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

extern "C" int lfgc_sample(const float* volume, const int32_t R[3], int64_t n, uint64_t seed, uint64_t sample_offset,
                           const int64_t* explicit_idx, float* raw_out, float* norm_out, float* gt_out, void* stream) {
    return lfgc_sample_stream(volume, R, n, seed, sample_offset, nullptr, 0, explicit_idx, raw_out, norm_out, gt_out,
                              stream);
}

extern "C" int lfgc_trilinear(const float* p, int64_t n, const float* volume, const int32_t R[3], const float min_bb[3],
                              const float max_bb[3], float* out, void* stream) {
    if (!R || !min_bb || !max_bb || n < 0 || !volume) return fail(LFGC_E_INVALID, "trilinear: bad arguments");
    if (n == 0) return LFGC_OK;
    if (!p || !out) return fail(LFGC_E_INVALID, "trilinear: null pointer");
    TriArgs A;
    A.p = p;
    A.f = volume;
    for (int a = 0; a < 3; ++a) {
        if (R[a] < 1) return fail(LFGC_E_INVALID, "trilinear: bad extent");
        A.R[a] = R[a];
        A.min_bb[a] = min_bb[a];
        A.max_bb[a] = max_bb[a];
    }
    A.n = n;
    A.out = out;
    trilinear_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}

extern "C" int lfgc_deviation_stats(const float* pred, const float* gt, int64_t n, double* acc, void* stream) {
    if (!pred || !gt || !acc || n < 0) return fail(LFGC_E_INVALID, "deviation_stats: bad arguments");
    if (n == 0) return LFGC_OK;
    int64_t blocks = (n + 1023) / 1024;
    const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 8;
    if (blocks > cap) blocks = cap;
    deviation_stats_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pred, gt, n, acc);
    LFGC_LAUNCH_OK();
    return LFGC_OK;
}
