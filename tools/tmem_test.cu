// TMEM as thread-private scratch: allocate 512 columns, every thread stores 64 words, reads them back in pieces.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_test tools/tmem_test.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__global__ void __launch_bounds__(256, 1) tmem_kernel(uint32_t *out, int *bad, long long *cycles) {
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t base = tmem_base;
    // this warp's lane quarter and column half
    const uint32_t my = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    // store 256 words per thread (64 x4 stores): value = tid * 1000 + word
    for (int w = 0; w < 256; w += 4)
        tmem_st4(my + w, tid * 1000 + w, tid * 1000 + w + 1, tid * 1000 + w + 2, tid * 1000 + w + 3);
    tmem_wait_st();
    int nbad = 0;
    long long t0 = clock64();
    for (int rep = 0; rep < 16; ++rep)
        for (int w = 0; w < 256; w += 4) {
            uint32_t a, b, c, d;
            tmem_ld4(my + w, a, b, c, d);
            tmem_wait_ld();
            if (a != (uint32_t)(tid * 1000 + w) || b != a + 1 || c != a + 2 || d != a + 3) ++nbad;
        }
    long long t1 = clock64();
    if (nbad) atomicAdd(bad, nbad);
    if (tid == 0) { cycles[blockIdx.x] = t1 - t0; out[blockIdx.x] = base; }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(base) : "memory");
}

int main() {
    uint32_t *out; int *bad; long long *cyc;
    cudaMalloc(&out, 148 * 4); cudaMalloc(&bad, 4); cudaMalloc(&cyc, 148 * 8); cudaMemset(bad, 0, 4);
    tmem_kernel<<<148, 256>>>(out, bad, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    int hbad = -1; long long hc[148]; uint32_t hb[148];
    cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost); cudaMemcpy(hb, out, sizeof(hb), cudaMemcpyDeviceToHost);
    printf("status %s  mismatches %d  base 0x%08x  cycles for 16x64 dependent x4 loads per thread: %lld (%.1f cyc per ld+wait)\n",
           cudaGetErrorString(e), hbad, hb[0], hc[0], hc[0] / (16.0 * 64));
    return 0;
}
