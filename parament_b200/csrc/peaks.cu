// peaks.cu -- pipe-peak microbenchmarks exported for the roofline denominators of bench.py:
// the path is bound by the FP64 pipe, for which MEASURED_PEAKS.json holds no figure, so the peak is measured
// on the same device in the same process.  Same kernels as tools/peaks.cu.
#include <cuda_runtime.h>
#include "../../include/parament.h"
#include "frag.cuh"

namespace {

template <int ILP>
__global__ void ffma_peak_kernel(float *out, int iters, float a, float b) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-6f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dfma_peak_kernel(double *out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-6 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma_peak_kernel(double *out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-6 + i; c1[i] = 0.5 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) pb::dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Legacy warp-level TF32 tensor path (mma.sync.m16n8k8, SASS HMMA.1688.F32.TF32): the only TF32 route for the
// register-resident small-dimension kernels (tcgen05.mma needs M >= 64 tiles in shared memory).
template <int ILP>
__global__ void tf32_mma_peak_kernel(float *out, int iters, float a, float b) {
    float c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-6f + i; c[i][1] = 0.5f * i; c[i][2] = 0.25f; c[i][3] = 0.125f; }
    const unsigned ua = __float_as_uint(a), ub = __float_as_uint(b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(ua), "r"(ua), "r"(ua), "r"(ua), "r"(ub), "r"(ub));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void copy_peak_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

}  // namespace

// kind: 0 FP32 FFMA TFLOP/s, 1 FP64 DFMA TFLOP/s, 2 FP64 DMMA (mma.sync.m8n8k4) TFLOP/s, 3 HBM copy GB/s (read+write),
// 4 TF32 mma.sync.m16n8k8 TFLOP/s (one TF32 product; the 3xTF32 split of the complex64 kernel executes three per fp32-grade product).
// Best of 5 timed launches after 2 warm-ups, CUDA events on the default stream of the current device.
extern "C" PARAMENT_API double Parament_measurePeak(int kind) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1.0;
    constexpr int ILP = 8;
    const int iters = 4096, threads = 256, blocks = sms * 4;   // 32 warps per SM
    const size_t copy_elems = (size_t)1 << 26;                  // 1 GiB each way
    void *buf = nullptr, *buf2 = nullptr;
    const size_t bytes = kind == 3 ? copy_elems * sizeof(double2) : (size_t)blocks * threads * sizeof(double);
    if (cudaMalloc(&buf, bytes) != cudaSuccess) return -1.0;
    if (kind == 3 && cudaMalloc(&buf2, bytes) != cudaSuccess) { cudaFree(buf); return -1.0; }
    if (kind == 3) cudaMemset(buf, 0, bytes);
    float best = 1e30f;
    for (int rep = 0; rep < 7; ++rep) {
        cudaEventRecord(e0);
        switch (kind) {
            case 0: ffma_peak_kernel<ILP><<<blocks, threads>>>((float *)buf, iters, 1.0001f, 0.5f); break;
            case 1: dfma_peak_kernel<ILP><<<blocks, threads>>>((double *)buf, iters, 1.0001, 0.5); break;
            case 2: dmma_peak_kernel<ILP><<<blocks, threads>>>((double *)buf, iters, 1.0001, 0.5); break;
            case 4: tf32_mma_peak_kernel<ILP><<<blocks, threads>>>((float *)buf, iters, 1.0001f, 0.5f); break;
            default: copy_peak_kernel<<<sms * 16, 256>>>((const double2 *)buf, (double2 *)buf2, copy_elems); break;
        }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    const bool ok = cudaGetLastError() == cudaSuccess;
    cudaFree(buf);
    if (buf2) cudaFree(buf2);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (!ok) return -1.0;
    const double nthr = (double)threads * blocks;
    switch (kind) {
        case 0:
        case 1: return 2.0 * ILP * iters * nthr / best * 1e-9;
        case 2: return 2.0 * 256 * ILP * iters * (nthr / 32) / best * 1e-9;
        case 4: return 2.0 * 1024 * ILP * iters * (nthr / 32) / best * 1e-9;
        default: return 2.0 * (double)bytes / best * 1e-6;
    }
}
