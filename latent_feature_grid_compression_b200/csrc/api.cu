// Library-level entry points and per-device caches.
#include "lfgc_common.cuh"

#include <atomic>
#include <math.h>
#include <stdlib.h>

namespace lfgc {

char* last_error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("LFGC_PDL");   // opt-in: measured on B200, graph-replayed step 82.0 us with PDL edges
        return e && e[0] == '1';              // against 79.9 us without (the graph already hides the launch latency)
    }();
    return on;
}

AdamCoef make_adam_coef(double beta1, double beta2, double eps, double grad_scale) {
    AdamCoef c;
    c.b1 = (float)beta1;
    c.b2 = (float)beta2;
    c.omb1 = (float)(1.0 - beta1);
    c.omb2 = (float)(1.0 - beta2);
    c.logb1 = (float)log(beta1);
    c.logb2 = (float)log(beta2);
    c.eps = (float)eps;
    c.gscale = (float)grad_scale;
    return c;
}

static int g_sm_count[64];
static int g_smem_optin[64];
static bool g_have[64];

static bool query_device(int& dev) {
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        cudaGetLastError();
        return false;
    }
    if (!g_have[dev]) {
        int sms = 0, smem = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        g_sm_count[dev] = sms;
        g_smem_optin[dev] = smem;
        g_have[dev] = true;
    }
    return true;
}

int sm_count() {
    int dev;
    return query_device(dev) ? g_sm_count[dev] : -1;
}

int max_smem_optin() {
    int dev;
    return query_device(dev) ? g_smem_optin[dev] : 0;
}

}  // namespace lfgc

extern "C" int lfgc_abi_version(void) { return LFGC_ABI_VERSION; }

// sizes of the structs that cross the ABI, in declaration order of include/lfgc.h: a binding checks its own layouts against
// these at load time (a silent mismatch would corrupt arguments)
extern "C" int lfgc_struct_sizes(size_t* out, int n) {
    const size_t sz[] = {sizeof(lfgc_wavelet_desc), sizeof(lfgc_model_desc), sizeof(lfgc_peer_announce), sizeof(lfgc_grid_step_args)};
    const int have = (int)(sizeof(sz) / sizeof(sz[0]));
    for (int i = 0; i < n && i < have; ++i) out[i] = sz[i];
    return have;
}
extern "C" const char* lfgc_last_error(void) { return lfgc::last_error_buffer(); }
extern "C" int lfgc_sm_count(void) { return lfgc::sm_count(); }
extern "C" long long lfgc_launch_count(void) { return lfgc::launch_count(); }
