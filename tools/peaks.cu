// Pipe-peak microbenchmarks for the roofline denominators (FFMA, DFMA, DMMA) on the box the bench
// runs on.  Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peaks peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void ffma_kernel(float *out, int iters, float a, float b) {
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
__global__ void dfma_kernel(double *out, int iters, double a, double b) {
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

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void dmma_kernel(double *out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-6 + i; c1[i] = 0.5 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k4 / m16n8k8 / m16n8k16 f64 shapes (sm_90+)
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void dmma1688_kernel(double *out, int iters, double av, double bv) {
    double c[ILP][4];
    double a[4] = {av, av + 1, av + 2, av + 3}, b[2] = {bv, bv + 1};
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-6 + i; c[i][1] = c[i][2] = c[i][3] = 0.5 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma16816_kernel(double *out, int iters, double av, double bv) {
    double c[ILP][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = av + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = bv + i;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-6 + i; c[i][1] = c[i][2] = c[i][3] = 0.5 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA and DFMA interleaved: do the two pipes overlap?
template <int ILP>
__global__ void mixed_kernel(double *out, int iters, double a, double b) {
    double c0[ILP], c1[ILP], f[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-6 + i; c1[i] = 0.5 * i; f[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            dmma884(c0[i], c1[i], a, b);
#pragma unroll
            for (int r = 0; r < 8; ++r) f[i] = fma(f[i], a, b);   // 8 DFMA = 8 FMA/lane, same work as 1 DMMA
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);
    void *buf; CK(cudaMalloc(&buf, sizeof(double) * sms * 64 * 1024));
    const int iters = 4096;
    for (int wps = 4; wps <= 32; wps *= 2) {       // warps per SM
        int threads = 256, blocks = sms * (wps * 32 / threads > 0 ? wps * 32 / threads : 1);
        if (wps * 32 < threads) { threads = wps * 32; blocks = sms; }
        double nthr = (double)threads * blocks;
        constexpr int ILP = 8;
        float t;
        t = time_ms([&] { ffma_kernel<ILP><<<blocks, threads>>>((float *)buf, iters, 1.0001f, 0.5f); }, 5);
        printf(", \"ffma_tflops_w%d\": %.2f", wps, 2.0 * ILP * iters * nthr / t * 1e-9);
        t = time_ms([&] { dfma_kernel<ILP><<<blocks, threads>>>((double *)buf, iters, 1.0001, 0.5); }, 5);
        printf(", \"dfma_tflops_w%d\": %.2f", wps, 2.0 * ILP * iters * nthr / t * 1e-9);
        t = time_ms([&] { dmma_kernel<ILP><<<blocks, threads>>>((double *)buf, iters, 1.0001, 0.5); }, 5);
        printf(", \"dmma884_tflops_w%d\": %.2f", wps, 2.0 * 256 * ILP * iters * (nthr / 32) / t * 1e-9);
        t = time_ms([&] { dmma1688_kernel<ILP><<<blocks, threads>>>((double *)buf, iters, 1.0001, 0.5); }, 5);
        printf(", \"dmma1688_tflops_w%d\": %.2f", wps, 2.0 * 1024 * ILP * iters * (nthr / 32) / t * 1e-9);
        t = time_ms([&] { dmma16816_kernel<ILP><<<blocks, threads>>>((double *)buf, iters, 1.0001, 0.5); }, 5);
        printf(", \"dmma16816_tflops_w%d\": %.2f", wps, 2.0 * 2048 * ILP * iters * (nthr / 32) / t * 1e-9);
        t = time_ms([&] { mixed_kernel<ILP><<<blocks, threads>>>((double *)buf, iters, 1.0001, 0.5); }, 5);
        printf(", \"mixed_dmma_dfma_tflops_w%d\": %.2f", wps, 2.0 * (256 + 256) * ILP * iters * (nthr / 32) / t * 1e-9);
    }
    // dependent-chain latency of DMMA (1 warp, ILP 1)
    {
        float t = time_ms([&] { dmma_kernel<1><<<1, 32>>>((double *)buf, 1 << 16, 1.0001, 0.5); }, 3);
        printf(", \"dmma884_latency_ns\": %.2f", t * 1e6 / (1 << 16));
        t = time_ms([&] { dfma_kernel<1><<<1, 32>>>((double *)buf, 1 << 16, 1.0001, 0.5); }, 3);
        printf(", \"dfma_latency_ns\": %.2f", t * 1e6 / (1 << 16));
        t = time_ms([&] { ffma_kernel<1><<<1, 32>>>((float *)buf, 1 << 16, 1.0001f, 0.5f); }, 3);
        printf(", \"ffma_latency_ns\": %.2f", t * 1e6 / (1 << 16));
    }
    printf("}\n");
    return 0;
}
