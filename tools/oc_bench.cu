// Microbenchmark of the shared-memory-resident complex 64x64x64 DMMA product in several variants (development aid).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I parament_b200/csrc -o tools/oc_bench tools/oc_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "frag.cuh"
using namespace pb;
constexpr int N = 64, P = 68, BUF = N * P;

// WM x WN warp tile, (64/WM)*(64/WN) warps.  mode: 0 = main loop only, 1 = + epilogue store to smem, 2 = + barrier per product
// EXTRA: doubles kept live across the main loop (register pressure like the real kernel's own-element arrays)
template <int WM, int WN, int UNROLL, int MODE, int EXTRA = 0>
__global__ void __launch_bounds__((64 / WM) * (64 / WN) * 32, 1) oc_variant(double2 *out, int reps) {
    extern __shared__ __align__(16) unsigned char raw[];
    double2 *b0 = reinterpret_cast<double2 *>(raw), *b1 = b0 + BUF, *b2 = b1 + BUF;
    constexpr int NT = (64 / WM) * (64 / WN) * 32;
    constexpr int MT = WM / 8, NTL = WN / 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, q = lane & 3;
    const int wm0 = (warp / (64 / WN)) * WM, wn0 = (warp % (64 / WN)) * WN;
    for (int e = tid; e < 3 * BUF; e += NT) b0[e] = make_double2(1e-3 * (e % 7), 1e-3 * (e % 5));
    __syncthreads();
    double2 *A = b0, *B = b1, *D = b2;
    double acc_keep = 0;
    double extra[EXTRA > 0 ? EXTRA : 1];
#pragma unroll
    for (int i = 0; i < EXTRA; ++i) extra[i] = b0[(tid * 7 + i) % BUF].x;
    for (int rep = 0; rep < reps; ++rep) {
        double cre[MT][NTL][2], cim[MT][NTL][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) { cre[mt][nt][0] = cre[mt][nt][1] = 0; cim[mt][nt][0] = cim[mt][nt][1] = 0; }
#pragma unroll UNROLL
        for (int kt = 0; kt < 16; ++kt) {
            double2 af[MT], bf[NTL];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) af[mt] = A[(wm0 + 8 * mt + gq) * P + 4 * kt + q];
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) bf[nt] = B[(4 * kt + q) * P + wn0 + 8 * nt + gq];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) {
                    dmma884(cre[mt][nt][0], cre[mt][nt][1], af[mt].x, bf[nt].x);
                    dmma884(cim[mt][nt][0], cim[mt][nt][1], af[mt].x, bf[nt].y);
                }
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) {
                    dmma884(cre[mt][nt][0], cre[mt][nt][1], af[mt].y, neg(bf[nt].y));
                    dmma884(cim[mt][nt][0], cim[mt][nt][1], af[mt].y, bf[nt].x);
                }
            }
        }
        if (EXTRA > 0) {   // consume the live values in the epilogue, as the real kernel does
#pragma unroll
            for (int i = 0; i < EXTRA; ++i) cre[(i / 4) % MT][(i / 2) % NTL][i % 2] += extra[i] * 1e-9;
        }
        if (MODE >= 1) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) {
                    const int r = wm0 + 8 * mt + gq, c = wn0 + 8 * nt + 2 * q;
                    D[r * P + c] = make_double2(cre[mt][nt][0] * 1e-3, cim[mt][nt][0] * 1e-3);
                    D[r * P + c + 1] = make_double2(cre[mt][nt][1] * 1e-3, cim[mt][nt][1] * 1e-3);
                }
        } else {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) acc_keep += cre[mt][nt][0] + cim[mt][nt][1];
        }
        if (MODE >= 2) { __syncthreads(); double2 *t = A; A = D; D = t; }
    }
    if (acc_keep == 123.456) out[0] = make_double2(acc_keep, 0);
    out[blockIdx.x * NT + tid] = b2[tid];
}

template <int WM, int WN, int UNROLL, int MODE, int EXTRA = 0>
void run(const char *name, double2 *out) {
    auto k = oc_variant<WM, WN, UNROLL, MODE, EXTRA>;
    const size_t smem = 3 * BUF * sizeof(double2);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int reps = 400, threads = (64 / WM) * (64 / WN) * 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148, threads, smem>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<148, threads, smem>>>(out, reps);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 148.0 * reps * 8.0 * 64 * 64 * 64;
    printf("%-44s threads %4d  %.3f ms  %.2f TFLOP/s  err=%s\n", name, threads, ms, flops / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double2 *out; cudaMalloc(&out, 148 * 1024 * sizeof(double2));
    run<32, 16, 2, 0>("8 warps 32x16 unroll2 loop only", out);
    run<32, 16, 4, 0>("8 warps 32x16 unroll4 loop only", out);
    run<32, 16, 1, 0>("8 warps 32x16 unroll1 loop only", out);
    run<32, 16, 2, 1>("8 warps 32x16 unroll2 + store", out);
    run<32, 16, 2, 2>("8 warps 32x16 unroll2 + store + barrier", out);
    run<16, 16, 2, 0>("16 warps 16x16 unroll2 loop only", out);
    run<16, 16, 2, 2>("16 warps 16x16 unroll2 + store + barrier", out);
    run<32, 32, 2, 0>("4 warps 32x32 unroll2 loop only", out);
    run<32, 32, 2, 2>("4 warps 32x32 unroll2 + store + barrier", out);
    run<16, 32, 2, 2>("8 warps 16x32 unroll2 + store + barrier", out);
    run<32, 16, 2, 2, 32>("8 warps 32x16 + store + barrier, 32 live doubles", out);
    run<32, 16, 2, 2, 64>("8 warps 32x16 + store + barrier, 64 live doubles", out);
    run<32, 16, 2, 2, 96>("8 warps 32x16 + store + barrier, 96 live doubles", out);
    return 0;
}
