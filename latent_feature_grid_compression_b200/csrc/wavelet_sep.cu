// Separable wavelet synthesis / adjoint for LARGE latent grids (mask-free levels): Feature_Grid_Model.decode_volume
// (model/Feature_Grid_Model.py:102-108) over _WaveletFilterNd.decode (wavelet_transform/Torch_Wavelet_Transform.py:91-104,
// grouped conv_transpose3d, stride 2, crop) and its autograd.
//
// wavelet.cu evaluates every output vertex as a direct sum over (n_taps/2)^3 x 8 coefficients: fine while the whole
// pyramid is a few hundred KB and the cost is launch latency, but 64 loads per output at C32/G64 (BASELINE.json configs[4]:
// 9.6 M coefficients) made the synthesis + adjoint 1.5 ms of a 1.9 ms step on B200 (r2 bench, 4 % of the HBM roofline of
// that work).  Here each level is three 1-D passes of n_taps MACs per output, two kernels per level and direction:
//
//   synthesis   S1  one CTA per (channel, input plane iz): the 8 band planes -> x pass -> y pass in shared memory ->
//                   Y[c][a][iz][oy][ox] (a = low / high along z)
//               S2  z pass, one thread per (channel, oy, ox) column -> out[c][oz][oy][ox]
//   adjoint     A1  z^T pass, one thread per column: g[c][oz][oy][ox] -> Y[c][a][iz][oy][ox]
//               A2  one CTA per (channel, iz): y^T pass -> x^T pass in shared memory -> the 8 band gradients
//
// All intermediates are CHANNEL-FIRST (the reference's coefficient layout), so every global access is a contiguous run;
// the z passes of the finest level turn 32 x 32 (position, channel) tiles in shared memory to read / write the
// channels-last grid the sample kernels use.
// HBM traffic per level ~ 4 x the level's output size instead of 64 L2 loads per vertex.
#include "wavelet_lines.cuh"

#include <stdlib.h>

namespace lfgc {
namespace wsep {

struct Level {
    int C, d[3], t[3], off[3], m_lo[3], n_m[3];
    float lo[LFGC_MAX_TAPS], hi[LFGC_MAX_TAPS];
    FastDiv by_d1, by_d2, by_nm2, by_plane;
    PassPlan sy, ay;           // the y passes: column (a, ox) fastest (contiguous global stores / conflict-free smem)
};

// ---- S1: x and y synthesis passes of one (channel, iz) plane ---------------------------------------------------------------
// low:  [C][d0][d1][d2] (coarser level's output, or coefficient tensor 0)     high: [C][7][d0][d1][d2]
// Y:    [C][2][d0][t1][t2]
template <int NT>
__global__ void __launch_bounds__(256) synth_xy_kernel(const __grid_constant__ Level L, const float* __restrict__ low,
                                                       const float* __restrict__ high, float* __restrict__ Y) {
    LFGC_PDL_PROLOGUE();
    extern __shared__ __align__(16) float sm[];
    const int d0 = L.d[0], d1 = L.d[1], d2 = L.d[2], t1 = L.t[1], t2 = L.t[2];
    const int iz = blockIdx.x, c = blockIdx.y;
    const int plane = d1 * d2;
    float* B = sm;                    // [8][d1][d2] band planes
    float* X = sm + 8 * plane;        // [4][d1][t2]
    float flo[NT], fhi[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        flo[i] = L.lo[i];
        fhi[i] = L.hi[i];
    }
    const size_t dvol = (size_t)d0 * plane;
    const float* lsrc = low + ((size_t)c * d0 + iz) * plane;
    const float* hsrc = high + ((size_t)c * 7 * d0 + iz) * plane;
    // B = [low plane | 7 high planes], one flat index space; eight independent loads in flight per thread (the plain
    // per-plane loops kept ~1 load in flight: the load phase was a chain of L2 round trips, long-scoreboard + barrier stalls)
    {
        const int total = 8 * plane;
        for (int base = threadIdx.x; base < total; base += 8 * blockDim.x) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * blockDim.x;
                v[u] = 0.0f;
                if (idx < total) {
                    const int k = fdiv(idx, L.by_plane);
                    const int i = idx - k * plane;
                    v[u] = __ldg(k == 0 ? lsrc + i : hsrc + (size_t)(k - 1) * dvol + i);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * blockDim.x;
                if (idx < total) B[idx] = v[u];
            }
        }
    }
    __syncthreads();
    // x pass: unit = (row = (ab, iy), pair m), m FASTEST: neighbouring threads write neighbouring ox (a row-fastest order
    // strides the shared-memory stores by t2 floats: 32-way bank conflicts at t2 = 64, measured +110 us per step at C32/G64)
    {
        const int n_m = L.n_m[2], nrows = 4 * d1;
        for (int u = threadIdx.x; u < nrows * n_m; u += blockDim.x) {
            const int row = fdiv(u, L.by_nm2);
            const int m = u - row * n_m + L.m_lo[2];
            const int ab = fdiv(row, L.by_d1);
            const int iy = row - ab * d1;
            synth_line<NT>(B + (2 * ab) * plane + iy * d2, B + (2 * ab + 1) * plane + iy * d2, 1, d2, X + row * t2, 1, t2,
                           L.off[2], m, m + 1, flo, fhi);
        }
    }
    __syncthreads();
    // y pass: unit = (range of pairs, column = (a, ox)), column fastest so that the global stores are contiguous in ox
    {
        float* out = Y + (size_t)c * 2 * d0 * t1 * t2 + (size_t)iz * t1 * t2;
        for (int u = threadIdx.x; u < L.sy.ncols * L.sy.groups; u += blockDim.x) {
            const int grp = fdiv(u, L.sy.by_ncols);
            const int col = u - grp * L.sy.ncols;
            const int a_ = col >= t2;
            const int ox = col - a_ * t2;
            const int r0 = grp * L.sy.per, r1 = min(r0 + L.sy.per, L.sy.n);
            synth_line<NT>(X + (2 * a_) * d1 * t2 + ox, X + (2 * a_ + 1) * d1 * t2 + ox, t2, d1,
                           out + (size_t)a_ * d0 * t1 * t2 + ox, t2, t1, L.off[1], L.m_lo[1] + r0, L.m_lo[1] + r1, flo, fhi);
        }
    }
}

// ---- S2: z synthesis pass, one thread per (channel, oy, ox) column ------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(256) synth_z_kernel(const __grid_constant__ Level L, const float* __restrict__ Y,
                                                      float* __restrict__ out) {
    LFGC_PDL_PROLOGUE();
    const int d0 = L.d[0], t0 = L.t[0], pl = L.t[1] * L.t[2];
    const int pos = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (pos >= pl) return;
    float flo[NT], fhi[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        flo[i] = L.lo[i];
        fhi[i] = L.hi[i];
    }
    const float* y0 = Y + (size_t)c * 2 * d0 * pl + pos;
    synth_line<NT>(y0, y0 + (size_t)d0 * pl, pl, d0, out + (size_t)c * t0 * pl + pos, pl, t0, L.off[0], L.m_lo[0],
                   L.m_lo[0] + L.n_m[0], flo, fhi);
}

// ---- A1: z^T adjoint pass, one thread per column -------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(256) adj_z_kernel(const __grid_constant__ Level L, const float* __restrict__ g,
                                                    float* __restrict__ Y) {
    LFGC_PDL_PROLOGUE();
    const int d0 = L.d[0], t0 = L.t[0], pl = L.t[1] * L.t[2];
    const int pos = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (pos >= pl) return;
    float flo[NT], fhi[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        flo[i] = L.lo[i];
        fhi[i] = L.hi[i];
    }
    float* y0 = Y + (size_t)c * 2 * d0 * pl + pos;
    adj_line<NT>(g + (size_t)c * t0 * pl + pos, pl, t0, L.off[0], y0, y0 + (size_t)d0 * pl, pl, 0, d0, flo, fhi);
}

// ---- A2: y^T and x^T adjoint passes of one (channel, iz) plane -------------------------------------------------------------------
// g_low: [C][d0][d1][d2] (gradient of the coarser level's output, or of coefficient tensor 0)    g_high: [C][7][d0][d1][d2]
template <int NT>
__global__ void __launch_bounds__(256) adj_yx_kernel(const __grid_constant__ Level L, const float* __restrict__ Y,
                                                     float* __restrict__ g_low, float* __restrict__ g_high) {
    LFGC_PDL_PROLOGUE();
    extern __shared__ __align__(16) float sm[];
    const int d0 = L.d[0], d1 = L.d[1], d2 = L.d[2], t1 = L.t[1], t2 = L.t[2];
    const int iz = blockIdx.x, c = blockIdx.y;
    float* Ys = sm;                      // [2][t1][t2]
    float* X = sm + 2 * t1 * t2;         // [4][d1][t2]
    float flo[NT], fhi[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        flo[i] = L.lo[i];
        fhi[i] = L.hi[i];
    }
    const int pl = t1 * t2;
    for (int a_ = 0; a_ < 2; ++a_) {
        const float* src = Y + ((size_t)(c * 2 + a_) * d0 + iz) * pl;
        if ((pl & 3) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0) {   // 128-bit loads, four in flight per thread
            const float4* s4 = reinterpret_cast<const float4*>(src);
            float4* d4 = reinterpret_cast<float4*>(Ys + a_ * pl);
            const int n4 = pl >> 2;
            for (int base = threadIdx.x; base < n4; base += 4 * blockDim.x) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = base + u * blockDim.x;
                    if (i < n4) v[u] = __ldg(s4 + i);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = base + u * blockDim.x;
                    if (i < n4) d4[i] = v[u];
                }
            }
        } else {
#pragma unroll 4
            for (int i = threadIdx.x; i < pl; i += blockDim.x) Ys[a_ * pl + i] = __ldg(src + i);
        }
    }
    __syncthreads();
    // y^T pass: unit = (range of iy, column = (a, ox)): Ys[a][:][ox] -> X[(a, b = 0 / 1)][iy][ox]
    for (int u = threadIdx.x; u < L.ay.ncols * L.ay.groups; u += blockDim.x) {
        const int grp = fdiv(u, L.ay.by_ncols);
        const int col = u - grp * L.ay.ncols;
        const int a_ = col >= t2;
        const int ox = col - a_ * t2;
        const int r0 = grp * L.ay.per, r1 = min(r0 + L.ay.per, L.ay.n);
        float* o0 = X + (2 * a_) * d1 * t2 + ox;
        adj_line<NT>(Ys + a_ * pl + ox, t2, t1, L.off[1], o0, o0 + d1 * t2, t2, r0, r1, flo, fhi);
    }
    __syncthreads();
    // x^T pass: unit = (row = (ab, iy), ix), ix fastest (see the x pass): X[ab][iy][:] -> bands (ab, c = 0 / 1)
    {
        const int nrows = 4 * d1;
        const size_t dvol = (size_t)d0 * d1 * d2;
        float* lowp = g_low + ((size_t)c * d0 + iz) * d1 * d2;
        float* highp = g_high + ((size_t)c * 7 * d0 + iz) * d1 * d2;
        for (int u = threadIdx.x; u < nrows * d2; u += blockDim.x) {
            const int row = fdiv(u, L.by_d2);
            const int ix = u - row * d2;
            const int ab = fdiv(row, L.by_d1);
            const int iy = row - ab * d1;
            float* o0 = (ab == 0 ? lowp : highp + (size_t)(2 * ab - 1) * dvol) + iy * d2;   // band (a, b, 0)
            float* o1 = highp + (size_t)(2 * ab) * dvol + iy * d2;                          // band (a, b, 1)
            adj_line<NT>(X + row * t2, 1, t2, L.off[2], o0, o1, 1, ix, ix + 1, flo, fhi);
        }
    }
}

// ---- finest level: the z passes read / write the CHANNELS-LAST grid directly ----------------------------------------------------
// A separate transpose of the 33.5 MB grid (and of its gradient) cost 39 + 33 us per step at C32/G64; here the 32 x 32
// (position, channel) tile is turned inside the z-pass kernels: compute with lane = position (contiguous Y accesses),
// touch the grid with lane = channel (contiguous channels-last accesses).
template <int NT>
__global__ void __launch_bounds__(1024) synth_z_cl_kernel(const __grid_constant__ Level L, const float* __restrict__ Y,
                                                          float* __restrict__ grid_cl, float* __restrict__ also_zero, int Cp) {
    LFGC_PDL_PROLOGUE();
    __shared__ float tile[2][32][33];
    const int d0 = L.d[0], t0 = L.t[0], pl = L.t[1] * L.t[2];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int pos = blockIdx.x * 32 + tx, c = blockIdx.y * 32 + ty;       // compute role: lane = position
    const int wpos = blockIdx.x * 32 + ty, wc = blockIdx.y * 32 + tx;      // write role:   lane = channel
    const bool live = pos < pl && c < L.C;
    float flo[NT], fhi[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        flo[i] = L.lo[i];
        fhi[i] = L.hi[i];
    }
    const float* y0 = Y + (size_t)(live ? c : 0) * 2 * d0 * pl + (live ? pos : 0);
    const float* y1 = y0 + (size_t)d0 * pl;
    // Sliding window over the input planes (each is used by NT / 2 consecutive m) with the NEXT plane's two loads issued
    // before this iteration's barriers and stores: the loop used to expose one global round trip per m (33 of them per CTA).
    constexpr int NW = NT / 2;
    float wl[NW], wh[NW];   // wl[a] = low input plane m - a
    // blockIdx.z splits the m range: 128 CTAs of 33 dependent iterations each left SMs idle and the chain long
    const int m_per = (L.n_m[0] + (int)gridDim.z - 1) / (int)gridDim.z;
    const int m_begin = L.m_lo[0] + (int)blockIdx.z * m_per;
    const int m_end = min(m_begin + m_per, L.m_lo[0] + L.n_m[0]);
    auto load_plane = [&](int i, float& l, float& h) {
        l = h = 0.0f;
        if (live && (unsigned)i < (unsigned)d0) {
            l = __ldg(y0 + (size_t)i * pl);
            h = __ldg(y1 + (size_t)i * pl);
        }
    };
#pragma unroll
    for (int a = 0; a < NW; ++a) load_plane(m_begin - a, wl[a], wh[a]);
    for (int m = m_begin; m < m_end; ++m) {
        float ln, hn;
        load_plane(m + 1 < m_end ? m + 1 : -1, ln, hn);
        float ev = 0.0f, od = 0.0f;
#pragma unroll
        for (int a = 0; a < NW; ++a) {
            ev = fmaf(wl[a], flo[2 * a], ev);
            ev = fmaf(wh[a], fhi[2 * a], ev);
            od = fmaf(wl[a], flo[2 * a + 1], od);
            od = fmaf(wh[a], fhi[2 * a + 1], od);
        }
#pragma unroll
        for (int a = NW - 1; a > 0; --a) {
            wl[a] = wl[a - 1];
            wh[a] = wh[a - 1];
        }
        wl[0] = ln;
        wh[0] = hn;
        tile[0][ty][tx] = ev;      // [channel][position]
        tile[1][ty][tx] = od;
        __syncthreads();
        const int oe = 2 * m - L.off[0];
        if (wpos < pl && wc < Cp) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int oz = oe + k;
                if ((unsigned)oz < (unsigned)t0) {
                    const size_t o = ((size_t)oz * pl + wpos) * Cp + wc;
                    grid_cl[o] = tile[k][tx][ty];
                    if (also_zero) also_zero[o] = 0.0f;
                }
            }
        }
        __syncthreads();
    }
}

template <int NT>
__global__ void __launch_bounds__(1024) adj_z_cl_kernel(const __grid_constant__ Level L, const float* __restrict__ g_cl, int Cp,
                                                        float* __restrict__ Y) {
    LFGC_PDL_PROLOGUE();
    __shared__ float tile[2][32][33];
    const int d0 = L.d[0], t0 = L.t[0], pl = L.t[1] * L.t[2];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int pos = blockIdx.x * 32 + tx, c = blockIdx.y * 32 + ty;       // compute role: lane = position
    const int rpos = blockIdx.x * 32 + ty, rc = blockIdx.y * 32 + tx;      // read role:    lane = channel
    const bool live = pos < pl && c < L.C;
    const bool rlive = rpos < pl && rc < L.C;
    float flo[NT], fhi[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        flo[i] = L.lo[i];
        fhi[i] = L.hi[i];
    }
    // window of the NT output planes q0 .. q0 + NT - 1 (q0 = 2 i - off) this thread's input position i sees; it slides by two
    float w[NT];
#pragma unroll
    for (int tt = 0; tt < NT; ++tt) w[tt] = 0.0f;
    float* y0 = Y + (size_t)(live ? c : 0) * 2 * d0 * pl + (live ? pos : 0);
    float* y1 = y0 + (size_t)d0 * pl;
    // planes below the first window position that are still inside it: load NT - 2 planes up front (q = -off .. -off + NT - 3)
    auto load_pair = [&](int i, float& a0, float& a1) {
        // the two planes entering the window of input position i: q = 2 i + NT - 2 - off and the next one
        const int qa = 2 * i + NT - 2 - L.off[0];
        a0 = (rlive && (unsigned)qa < (unsigned)t0) ? __ldg(g_cl + ((size_t)qa * pl + rpos) * Cp + rc) : 0.0f;
        a1 = (rlive && (unsigned)(qa + 1) < (unsigned)t0) ? __ldg(g_cl + ((size_t)(qa + 1) * pl + rpos) * Cp + rc) : 0.0f;
    };
    // blockIdx.z splits the input positions; every chunk first fills its window (NT - 2 planes)
    const int i_per = (d0 + (int)gridDim.z - 1) / (int)gridDim.z;
    const int i_begin = (int)blockIdx.z * i_per, i_end = min(i_begin + i_per, d0);
    float c0, c1;
    load_pair(i_begin - (NT - 2) / 2, c0, c1);
    for (int i = i_begin - (NT - 2) / 2; i < i_end; ++i) {
        float n0 = 0.0f, n1 = 0.0f;
        if (i + 1 < i_end) load_pair(i + 1, n0, n1);   // next iteration's planes: in flight across the barriers below
        tile[0][ty][tx] = c0;
        tile[1][ty][tx] = c1;
        c0 = n0;
        c1 = n1;
        __syncthreads();
#pragma unroll
        for (int tt = 0; tt + 2 < NT; ++tt) w[tt] = w[tt + 2];
        w[NT - 2] = tile[0][tx][ty];      // [position][channel] as stored by the read role
        w[NT - 1] = tile[1][tx][ty];
        __syncthreads();
        if (i >= i_begin && live) {
            float lo = 0.0f, hi = 0.0f;
#pragma unroll
            for (int tt = 0; tt < NT; ++tt) {
                lo = fmaf(w[tt], flo[tt], lo);
                hi = fmaf(w[tt], fhi[tt], hi);
            }
            y0[(size_t)i * pl] = lo;
            y1[(size_t)i * pl] = hi;
        }
    }
}

static void fill_level(Level& L, const lfgc_wavelet_desc* w, int l) {
    L.C = w->C;
    for (int i = 0; i < LFGC_MAX_TAPS; ++i) {
        L.lo[i] = i < w->n_taps ? w->rec_lo[i] : 0.0f;
        L.hi[i] = i < w->n_taps ? w->rec_hi[i] : 0.0f;
    }
    for (int a = 0; a < 3; ++a) {
        L.d[a] = w->dims[l][a];
        L.t[a] = w->target[l][a];
        L.off[a] = (2 * L.d[a] + w->n_taps - 2 - L.t[a]) / 2;   // floor(delta / 2), Torch_Wavelet_Transform.py:71
        L.m_lo[a] = L.off[a] >> 1;
        L.n_m[a] = ((L.off[a] + L.t[a] - 1) >> 1) - L.m_lo[a] + 1;
    }
    L.by_d1 = make_fastdiv((unsigned)L.d[1]);
    L.by_d2 = make_fastdiv((unsigned)L.d[2]);
    L.by_nm2 = make_fastdiv((unsigned)L.n_m[2]);
    L.by_plane = make_fastdiv((unsigned)(L.d[1] * L.d[2]));
    // one position per unit: these passes are latency-bound, more independent units hide it better
    L.sy = make_plan(2 * L.t[2], L.n_m[1], 1 << 24);
    L.ay = make_plan(2 * L.t[2], L.d[1], 1 << 24);
}

static size_t smem_s1(const Level& L) { return ((size_t)8 * L.d[1] * L.d[2] + (size_t)4 * L.d[1] * L.t[2]) * sizeof(float); }
static size_t smem_a2(const Level& L) { return ((size_t)2 * L.t[1] * L.t[2] + (size_t)4 * L.d[1] * L.t[2]) * sizeof(float); }

}  // namespace wsep

// ---- what wavelet.cu's entry points call -----------------------------------------------------------------------------------------

// floats of scratch the separable path needs (0: the descriptor is not covered)
size_t wavelet_sep_scratch_elems(const lfgc_wavelet_desc* w) {
    if (w->n_coeff < 2 || (w->n_taps != 2 && w->n_taps != 4)) return 0;
    size_t ybuf = 0, inter = 0;
    const int cap = max_smem_optin() > 0 ? max_smem_optin() : 232448;
    for (int l = 1; l < w->n_coeff; ++l) {
        wsep::Level L;
        wsep::fill_level(L, w, l);
        if (wsep::smem_s1(L) > (size_t)cap || wsep::smem_a2(L) > (size_t)cap) return 0;
        if (L.d[0] > 65535 || w->C > 65535) return 0;
        const size_t y = (size_t)w->C * 2 * L.d[0] * L.t[1] * L.t[2];
        const size_t o = (size_t)w->C * L.t[0] * L.t[1] * L.t[2];
        if (y > ybuf) ybuf = y;
        if (o > inter) inter = o;
    }
    return ybuf + 2 * inter + 64;
}

// The separable path pays two launches per level plus a transpose: worth it once the finest level is big enough that the
// direct sums are work- rather than latency-bound (measured crossover on B200 between G = 15 and G = 33 at C = 16..32).
bool wavelet_sep_preferred(const lfgc_wavelet_desc* w) {
    if (wavelet_sep_scratch_elems(w) == 0) return false;
    if (const char* e = getenv("LFGC_WAVELET_SEP")) {   // "0": never, "1": whenever the descriptor is covered (tests)
        if (e[0] == '0') return false;
        if (e[0] == '1') return true;
    }
    const int l = w->n_coeff - 1;
    const double out = (double)w->C * w->target[l][0] * w->target[l][1] * w->target[l][2];
    return out >= 1.5e6;
}

template <int NT>
static int sep_fwd(const lfgc_wavelet_desc* w, const float* const* coeff, float* scratch, float* grid_cl, int Cp,
                   float* also_zero, cudaStream_t st) {
    using namespace wsep;
    size_t ybuf = 0, inter = 0;
    for (int l = 1; l < w->n_coeff; ++l) {
        const size_t y = (size_t)w->C * 2 * w->dims[l][0] * w->target[l][1] * w->target[l][2];
        const size_t o = (size_t)w->C * w->target[l][0] * w->target[l][1] * w->target[l][2];
        if (y > ybuf) ybuf = y;
        if (o > inter) inter = o;
    }
    float* Y = scratch;
    float* buf[2] = {scratch + ybuf, scratch + ybuf + inter};
    const float* low = coeff[0];
    for (int l = 1; l < w->n_coeff; ++l) {
        Level L;
        fill_level(L, w, l);
        auto k1 = synth_xy_kernel<NT>;
        const size_t sm1 = smem_s1(L);
        LFGC_CUDA_OK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
        (void)launch_pdl(k1, dim3((unsigned)L.d[0], (unsigned)L.C), dim3(256), sm1, st, L, low, coeff[l], Y);
        LFGC_LAUNCH_OK();
        const int pl = L.t[1] * L.t[2];
        if (l == w->n_coeff - 1) {   // finest level: straight into the channels-last grid
            (void)launch_pdl(synth_z_cl_kernel<NT>, dim3((unsigned)((pl + 31) / 32), (unsigned)((Cp + 31) / 32), 2u), dim3(1024), (size_t)0,
                             st, L, (const float*)Y, grid_cl, also_zero, Cp);
            LFGC_LAUNCH_OK();
            break;
        }
        float* out = buf[l & 1];
        (void)launch_pdl(synth_z_kernel<NT>, dim3((unsigned)((pl + 255) / 256), (unsigned)L.C), dim3(256), (size_t)0, st, L,
                         (const float*)Y, out);
        LFGC_LAUNCH_OK();
        low = out;
    }
    return LFGC_OK;
}

template <int NT>
static int sep_bwd(const lfgc_wavelet_desc* w, const float* grad_grid_cl, int Cp, float* scratch, float* const* grad_coeff,
                   cudaStream_t st) {
    using namespace wsep;
    size_t ybuf = 0, inter = 0;
    for (int l = 1; l < w->n_coeff; ++l) {
        const size_t y = (size_t)w->C * 2 * w->dims[l][0] * w->target[l][1] * w->target[l][2];
        const size_t o = (size_t)w->C * w->target[l][0] * w->target[l][1] * w->target[l][2];
        if (y > ybuf) ybuf = y;
        if (o > inter) inter = o;
    }
    float* Y = scratch;
    float* buf[2] = {scratch + ybuf, scratch + ybuf + inter};
    const int last = w->n_coeff - 1;
    const float* g = nullptr;
    for (int l = last; l >= 1; --l) {
        Level L;
        fill_level(L, w, l);
        const int pl = L.t[1] * L.t[2];
        if (l == last) {   // finest level: straight from the channels-last gradient
            (void)launch_pdl(adj_z_cl_kernel<NT>, dim3((unsigned)((pl + 31) / 32), (unsigned)((L.C + 31) / 32), 2u), dim3(1024), (size_t)0,
                             st, L, grad_grid_cl, Cp, Y);
        } else {
            (void)launch_pdl(adj_z_kernel<NT>, dim3((unsigned)((pl + 255) / 256), (unsigned)L.C), dim3(256), (size_t)0, st, L, g, Y);
        }
        LFGC_LAUNCH_OK();
        float* g_low = l == 1 ? grad_coeff[0] : buf[(l - 1) & 1];
        auto k2 = adj_yx_kernel<NT>;
        const size_t sm2 = smem_a2(L);
        LFGC_CUDA_OK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        (void)launch_pdl(k2, dim3((unsigned)L.d[0], (unsigned)L.C), dim3(256), sm2, st, L, (const float*)Y, g_low, grad_coeff[l]);
        LFGC_LAUNCH_OK();
        g = g_low;
    }
    return LFGC_OK;
}

int wavelet_sep_fwd(const lfgc_wavelet_desc* w, const float* const* coeff, float* scratch, float* grid_cl, int Cp,
                    float* also_zero, cudaStream_t st) {
    return w->n_taps == 2 ? sep_fwd<2>(w, coeff, scratch, grid_cl, Cp, also_zero, st)
                          : sep_fwd<4>(w, coeff, scratch, grid_cl, Cp, also_zero, st);
}
int wavelet_sep_bwd(const lfgc_wavelet_desc* w, const float* grad_grid_cl, int Cp, float* scratch, float* const* grad_coeff,
                    cudaStream_t st) {
    return w->n_taps == 2 ? sep_bwd<2>(w, grad_grid_cl, Cp, scratch, grad_coeff, st)
                          : sep_bwd<4>(w, grad_grid_cl, Cp, scratch, grad_coeff, st);
}

}  // namespace lfgc
