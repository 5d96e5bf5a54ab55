// Microbenchmark of the register-resident 16x16 complex DMMA product as the warp kernel issues it (development aid):
// what fraction of the FP64 tensor peak do 2 (or 3) warps per scheduler reach with (a) products only, (b) + layout shuffles,
// (c) + FP64 scalar work between products.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I parament_b200/csrc -o tools/k1_bench tools/k1_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "frag.cuh"
using namespace pb;

template <int MODE, int OCC>
__global__ void __launch_bounds__(128, OCC) k(double2 *out, int reps, double eps, double c) {
    const int lane = threadIdx.x & 31;
    AccFrag<2> Q, E;
    set_identity<2>(Q, lane);
    set_identity<2>(E, lane);
#pragma unroll
    for (int e = 0; e < 8; ++e) { (&E.re[0][0][0])[e] *= eps; (&E.im[0][0][0])[e] = eps * 0.5 * (lane & 3); }
    for (int r = 0; r < reps; ++r) {
        BFrag<2> Et;
        if (MODE == 1 || MODE == 3) {
            acc_to_bfrag<2>(Et, E, lane);
#pragma unroll
            for (int e = 0; e < 8; ++e) (&Et.nim[0][0])[e] = neg((&Et.im[0][0])[e]);
        } else {
            transpose_as_bfrag<2>(Et, E);
        }
        AccFrag<2> Qn = Q;
        cmma<2>(Qn, Q, Et);
        Q = Qn;
        if (MODE >= 2) {   // 48 FP64 scalar instructions depending on the product
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                (&E.re[0][0][0])[e] = fma(c, (&Q.im[0][0][0])[e], fma(-c, (&Q.re[0][0][0])[e], (&E.re[0][0][0])[e]));
                (&E.im[0][0][0])[e] = fma(c, (&Q.re[0][0][0])[e], fma(c, (&Q.im[0][0][0])[e], (&E.im[0][0][0])[e]));
                (&E.re[0][0][0])[e] = fma(c, (&E.im[0][0][0])[e], (&E.re[0][0][0])[e]);
            }
        }
    }
    store_acc<2>(Q, out + (size_t)(blockIdx.x * 4 + (threadIdx.x >> 5)) * 256, 16, lane);
}

template <int MODE, int OCC>
void run(const char *name, double2 *out, double peak) {
    const int reps = 20000;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE, OCC><<<148 * OCC, 128>>>(out, 100, 1e-9, 1e-12);
    cudaEventRecord(a);
    k<MODE, OCC><<<148 * OCC, 128>>>(out, reps, 1e-9, 1e-12);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double tf = 148.0 * OCC * 4 * reps * 32768.0 / (ms * 1e-3) / 1e12;
    printf("%-44s occ %d  %.3f ms  %.2f TFLOP/s  (%.1f %% of %.1f)\n", name, OCC, ms, tf, 100 * tf / peak, peak);
}

int main() {
    double2 *out;
    cudaMalloc(&out, 148 * 8 * 4 * 256 * sizeof(double2));
    const double peak = 36.9;
    run<0, 2>("products only", out, peak);
    run<0, 3>("products only", out, peak);
    run<0, 4>("products only", out, peak);
    run<1, 2>("products + layout shuffle", out, peak);
    run<1, 3>("products + layout shuffle", out, peak);
    run<2, 2>("products + 40 DFMA", out, peak);
    run<2, 3>("products + 40 DFMA", out, peak);
    run<2, 4>("products + 40 DFMA", out, peak);
    run<3, 2>("products + shuffle + 40 DFMA", out, peak);
    run<3, 3>("products + shuffle + 40 DFMA", out, peak);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
