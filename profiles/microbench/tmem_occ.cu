#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 4) k_tmem(int* out) {
    __shared__ unsigned slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"((unsigned)__cvta_generic_to_shared(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    __syncthreads();
    unsigned base = slot;
    if (threadIdx.x == 0) out[blockIdx.x] = (int)base;
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base));
}
__global__ void __launch_bounds__(128, 4) k_plain(int* out) { if (threadIdx.x == 0) out[blockIdx.x] = 1; }
int main() {
    for (int smem : {0, 16384, 66304}) {
        int o1 = -1, o2 = -1;
        cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, k_tmem, 128, smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, k_plain, 128, smem);
        printf("dyn smem %6d: occupancy tmem-kernel %d, plain kernel %d\n", smem, o1, o2);
    }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_tmem);
    printf("k_tmem: regs %d static smem %zu maxdyn %d\n", fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes);
    int* out; cudaMalloc(&out, 4096 * 4);
    k_tmem<<<148 * 4, 128>>>(out); cudaDeviceSynchronize();
    printf("run: %s\n", cudaGetErrorString(cudaGetLastError()));
    int h[16]; cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 8; ++i) printf("base[%d]=0x%x ", i, h[i]); printf("\n");
}
